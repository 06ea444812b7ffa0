// Depthwise kxk conv + LayerNorm(C, eps 1e-6) + AdaLN (or affine LN), NHWC fp32 in, fp32 and/or bf16 planes out.
// Reference: ConvNeXtBlockAdaLN.forward, lvae/models/common.py:145-152 (conv_dw -> permute -> norm ->
// x*(1+scale)+shift); qresvae MyConvNeXtBlock (affine LayerNorm, no AdaLN) qresvae/model.py:163-182.
//
// HBM-bound stage (algorithmic bytes per position: read 4C, write 4C fp32 or 2C per bf16 plane).  One CTA of
// 8 warps owns an 8 x 8 tile of output pixels of one image, all C channels:
//   * channels are processed in chunks of 64; the (8+k-1)^2 halo of a chunk is one TMA box load
//     (cp.async.bulk.tensor.4d over the [C, W, H, B] view; coordinates outside the image are zero-filled by the
//     hardware = the conv's zero padding), double-buffered on mbarriers so chunk j+1 streams in while chunk j is
//     convolved -- no fill loop, no index arithmetic, every input element is read once per tile;
//   * warp w convolves output row w: lane l owns channels {64 j + 2 l, 64 j + 2 l + 1}, a sliding window of the
//     shared-memory row feeds the 8 pixels, results stay in registers (8 pixels x 2 channels x C/64 chunks);
//   * LayerNorm is then warp-local: two-pass (mean, centred second moment) warp-shuffle reductions in fp32,
//     followed by the modulation and the split into bf16 planes for the tensor-core GEMM that consumes it;
//   * wide layers (C > 512, the rd model's 640 / 768) run as a 2-CTA thread-block cluster: each CTA convolves half
//     of the channels of the same pixel tile and the two exchange their per-pixel LayerNorm partial sums through
//     distributed shared memory (always added in rank order, so both CTAs -- and every launch -- agree bit for bit).
#include "common.cuh"
#include <cuda_bf16.h>
#include <cuda.h>
#include <cooperative_groups.h>
#include <mutex>
#include <stdlib.h>

namespace lvae {

constexpr int DW_T = 8;                 // output tile edge
constexpr int DW_CH = 64;               // channels per chunk
constexpr int DW_NBUF = 2;              // halo buffers per CTA (chunk j+1 streams in while chunk j is convolved)

__device__ __forceinline__ uint32_t dw_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dw_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void dw_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void dw_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void dw_tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// packed fp32x2 FMA (sm_100): both lanes are IEEE fma, i.e. bit-identical to two fmaf() at half the issue slots
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d)
      : "l"(*reinterpret_cast<const uint64_t*>(&a)), "l"(*reinterpret_cast<const uint64_t*>(&b)),
        "l"(*reinterpret_cast<const uint64_t*>(&c)));
  return *reinterpret_cast<float2*>(&d);
}

template <int NJ, int KS, int CL>
__global__ void __launch_bounds__(256) dwln_kernel(
    const __grid_constant__ CUtensorMap x_map, const float* __restrict__ dw_w, const float* __restrict__ dw_b,
    const float* __restrict__ ada, int64_t ada_stride, int64_t ada_off,
    const float* __restrict__ ln_w, const float* __restrict__ ln_b,
    float* __restrict__ y, __nv_bfloat16* __restrict__ y0, __nv_bfloat16* __restrict__ y1, __nv_bfloat16* __restrict__ y2,
    int f16, int H, int W, int tiles_x, int tiles_y) {
  constexpr int C = NJ * DW_CH * CL, PAD = (KS - 1) / 2, HT = DW_T + KS - 1;   // halo tile edge
  constexpr int CHUNK_FLOATS = HT * HT * DW_CH;
  extern __shared__ __align__(128) float dw_smem[];                          // [DW_NBUF][HT][HT][64]
  __shared__ __align__(8) uint64_t dw_bar[2];
  __shared__ float ln_part[2][8][DW_T];                                      // [pass][warp][pixel] partial sums (CL > 1)
  const int tid = threadIdx.x, lane = tid & 31, wrow = tid >> 5;
  const int crank = (CL > 1) ? (int)(blockIdx.x % CL) : 0;                   // cluster dims (CL,1,1): rank == blockIdx.x % CL
  const int cbase = crank * NJ * DW_CH;                                      // first channel of this CTA
  int t = blockIdx.x / CL;
  const int tx = t % tiles_x; t /= tiles_x;
  const int ty = t % tiles_y; const int b = t / tiles_y;
  const int h0 = ty * DW_T, w0 = tx * DW_T;

  if (tid == 0) {
    dw_mbar_init(dw_smem_u32(&dw_bar[0]), 1);
    dw_mbar_init(dw_smem_u32(&dw_bar[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto load_chunk = [&](int j, int buf) {                                    // one thread
    const uint32_t bar = dw_smem_u32(&dw_bar[buf]);
    dw_mbar_expect_tx(bar, (uint32_t)(CHUNK_FLOATS * 4));
    dw_tma_load_4d(dw_smem_u32(dw_smem + buf * CHUNK_FLOATS), &x_map, bar, cbase + j * DW_CH, w0 - PAD, h0 - PAD, b);
  };

  float2 res[NJ][DW_T];
  if (tid == 0) load_chunk(0, 0);
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    // buffer (j+1)&1 was last read in iteration j-1, which ended with __syncthreads()
    if (tid == 0 && j + 1 < NJ) load_chunk(j + 1, (j + 1) & 1);
    dw_mbar_wait(dw_smem_u32(&dw_bar[j & 1]), (uint32_t)((j >> 1) & 1));
    const float* tile = dw_smem + (j & 1) * CHUNK_FLOATS;
    const int c = cbase + j * DW_CH + lane * 2;
    const float2 bias = __ldg(reinterpret_cast<const float2*>(dw_b + c));
    float2 acc[DW_T];
#pragma unroll
    for (int s = 0; s < DW_T; ++s) acc[s] = bias;
#pragma unroll
    for (int ky = 0; ky < KS; ++ky) {
      const float* row = tile + ((wrow + ky) * HT) * DW_CH + lane * 2;
      float2 xv[HT];
#pragma unroll
      for (int i = 0; i < HT; ++i) xv[i] = *reinterpret_cast<const float2*>(row + i * DW_CH);
#pragma unroll
      for (int kx = 0; kx < KS; ++kx) {
        const float2 wv = __ldg(reinterpret_cast<const float2*>(dw_w + (ky * KS + kx) * C + c));
#pragma unroll
        for (int s = 0; s < DW_T; ++s) acc[s] = ffma2(xv[s + kx], wv, acc[s]);
      }
    }
#pragma unroll
    for (int s = 0; s < DW_T; ++s) res[j][s] = acc[s];
    __syncthreads();                       // everyone is done with this buffer before chunk j+2 overwrites it
  }

  const int h = h0 + wrow;
  if (CL == 1 && h >= H) return;           // warp-uniform (cluster CTAs stay for the exchanges below)
  namespace cg = cooperative_groups;
  float mean[DW_T], rstd[DW_T];
  float part[DW_T];
#pragma unroll
  for (int s = 0; s < DW_T; ++s) {
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) sum += res[j][s].x + res[j][s].y;
    part[s] = warp_sum(sum);
  }
  if (CL > 1) {
    if (lane == 0) {
#pragma unroll
      for (int s = 0; s < DW_T; ++s) ln_part[0][wrow][s] = part[s];
    }
    cg::this_cluster().sync();
#pragma unroll
    for (int s = 0; s < DW_T; ++s) {
      float tot = 0.f;
      for (int r = 0; r < CL; ++r) tot += cg::this_cluster().map_shared_rank(&ln_part[0][wrow][s], r)[0];
      part[s] = tot;
    }
  }
#pragma unroll
  for (int s = 0; s < DW_T; ++s) {
    mean[s] = part[s] * (1.0f / C);
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const float dx = res[j][s].x - mean[s], dy = res[j][s].y - mean[s];
      sq = fmaf(dx, dx, sq); sq = fmaf(dy, dy, sq);
    }
    part[s] = warp_sum(sq);
  }
  if (CL > 1) {
    if (lane == 0) {
#pragma unroll
      for (int s = 0; s < DW_T; ++s) ln_part[1][wrow][s] = part[s];
    }
    cg::this_cluster().sync();
#pragma unroll
    for (int s = 0; s < DW_T; ++s) {
      float tot = 0.f;
      for (int r = 0; r < CL; ++r) tot += cg::this_cluster().map_shared_rank(&ln_part[1][wrow][s], r)[0];
      part[s] = tot;
    }
    cg::this_cluster().sync();               // nobody exits while a peer may still read its shared memory
    if (h >= H) return;
  }
#pragma unroll
  for (int s = 0; s < DW_T; ++s) rstd[s] = 1.0f / sqrtf(part[s] * (1.0f / C) + 1e-6f);
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int c = cbase + j * DW_CH + lane * 2;
    float2 mul, add;                       // v * mul + add with mul = (1 + scale) | gamma, add = shift | beta
    if (ln_w != nullptr) {
      mul = __ldg(reinterpret_cast<const float2*>(ln_w + c));
      add = __ldg(reinterpret_cast<const float2*>(ln_b + c));
    } else {
      const float* e = ada + (int64_t)b * ada_stride + ada_off + c;
      add = __ldg(reinterpret_cast<const float2*>(e));
      const float2 sc = __ldg(reinterpret_cast<const float2*>(e + C));
      mul = make_float2(__fadd_rn(1.0f, sc.x), __fadd_rn(1.0f, sc.y));
    }
#pragma unroll
    for (int s = 0; s < DW_T; ++s) {
      const int w = w0 + s;
      if (w >= W) break;                   // warp-uniform
      float2 v;
      v.x = __fadd_rn(__fmul_rn(__fmul_rn(res[j][s].x - mean[s], rstd[s]), mul.x), add.x);
      v.y = __fadd_rn(__fmul_rn(__fmul_rn(res[j][s].y - mean[s], rstd[s]), mul.y), add.y);
      const int64_t o = (((int64_t)b * H + h) * W + w) * C + c;
      if (y != nullptr) *reinterpret_cast<float2*>(y + o) = v;
      if (y0 != nullptr) {
        // A operand of the tensor-core fc1 GEMM: p0 = rn16(v), p1 = rn16(v - p0), p2 = rn16(v - p0 - p1)
        *reinterpret_cast<uint32_t*>(y0 + o) = split_next(v, f16 != 0);
        if (y1 != nullptr) {
          *reinterpret_cast<uint32_t*>(y1 + o) = split_next(v, f16 != 0);
          if (y2 != nullptr) *reinterpret_cast<uint32_t*>(y2 + o) = split_next(v, f16 != 0);
        }
      }
    }
  }
}

typedef CUresult (*DwEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static DwEncodeTiledFn dw_encode_fn() {
  static DwEncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<DwEncodeTiledFn>(ptr);
  });
  return fn;
}

template <int NJ, int KS, int CL>
static int launch_dwln(const float* x, const float* dw_w, const float* dw_b, const float* ada,
                       int64_t ada_stride, int64_t ada_off, const float* ln_w, const float* ln_b,
                       float* y, __nv_bfloat16* y0, __nv_bfloat16* y1, __nv_bfloat16* y2, int f16,
                       int B, int H, int W, cudaStream_t stream) {
  constexpr int HT = DW_T + KS - 1, C = NJ * DW_CH * CL;
  constexpr int smem = DW_NBUF * HT * HT * DW_CH * 4;
  static bool configured = false;
  if (!configured) {
    LVAE_CUDA_CALL(cudaFuncSetAttribute(dwln_kernel<NJ, KS, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  DwEncodeTiledFn enc = dw_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return LVAE_E_UNSUPPORTED; }
  // NHWC fp32 viewed as a 4-D tensor (C, W, H, B); box = (64 channels, HT, HT, 1); out-of-image coordinates -> 0
  CUtensorMap map;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  cuuint32_t box[4] = {DW_CH, HT, HT, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (dwconv input) failed: %d", (int)r); return LVAE_E_BADARG; }
  const int tiles_x = (W + DW_T - 1) / DW_T, tiles_y = (H + DW_T - 1) / DW_T;
  const int64_t blocks = (int64_t)B * tiles_x * tiles_y;
  if (CL == 1) {
    dwln_kernel<NJ, KS, CL><<<(unsigned)blocks, 256, smem, stream>>>(
        map, dw_w, dw_b, ada, ada_stride, ada_off, ln_w, ln_b, y, y0, y1, y2, f16, H, W, tiles_x, tiles_y);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(blocks * CL)); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    LVAE_CUDA_CALL(cudaLaunchKernelEx(&cfg, dwln_kernel<NJ, KS, CL>, map, dw_w, dw_b, ada, ada_stride, ada_off, ln_w, ln_b,
                                      y, y0, y1, y2, f16, H, W, tiles_x, tiles_y));
  }
  LVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

template <int NJ, int CL>
static int dispatch_k(int k, const float* x, const float* dw_w, const float* dw_b, const float* ada,
                      int64_t ada_stride, int64_t ada_off, const float* ln_w, const float* ln_b,
                      float* y, __nv_bfloat16* y0, __nv_bfloat16* y1, __nv_bfloat16* y2, int f16,
                      int B, int H, int W, cudaStream_t stream) {
  switch (k) {
    case 1: return launch_dwln<NJ, 1, CL>(x, dw_w, dw_b, ada, ada_stride, ada_off, ln_w, ln_b, y, y0, y1, y2, f16, B, H, W, stream);
    case 3: return launch_dwln<NJ, 3, CL>(x, dw_w, dw_b, ada, ada_stride, ada_off, ln_w, ln_b, y, y0, y1, y2, f16, B, H, W, stream);
    case 5: return launch_dwln<NJ, 5, CL>(x, dw_w, dw_b, ada, ada_stride, ada_off, ln_w, ln_b, y, y0, y1, y2, f16, B, H, W, stream);
    case 7: return launch_dwln<NJ, 7, CL>(x, dw_w, dw_b, ada, ada_stride, ada_off, ln_w, ln_b, y, y0, y1, y2, f16, B, H, W, stream);
    default: set_error("dwconv kernel size %d unsupported", k); return LVAE_E_UNSUPPORTED;
  }
}

}  // namespace lvae

static int dwln_dispatch(const float* x, const float* dw_w, const float* dw_b,
                         const float* ada, int64_t ada_stride, int64_t ada_off,
                         const float* ln_w, const float* ln_b, float* y, void* y0, void* y1, void* y2, int f16,
                         int B, int H, int W, int C, int k, void* stream) {
  using namespace lvae;
  LVAE_CHECK_ARG(x && dw_w && dw_b && (y || y0) && (ada || ln_w));
  LVAE_CHECK_ARG(B > 0 && H > 0 && W > 0 && C > 0 && C % 64 == 0);
  LVAE_CHECK_ARG(ln_w != nullptr || (ada_off % 2 == 0 && ada_stride % 2 == 0));      // float2 loads of shift / scale
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* p0 = (__nv_bfloat16*)y0; __nv_bfloat16* p1 = (__nv_bfloat16*)y1; __nv_bfloat16* p2 = (__nv_bfloat16*)y2;
#define LVAE_DWLN_CASE(nj) case nj: return dispatch_k<nj, 1>(k, x, dw_w, dw_b, ada, ada_stride, ada_off, ln_w, ln_b, y, p0, p1, p2, f16, B, H, W, st);
  switch (C / 64) {
    LVAE_DWLN_CASE(1) LVAE_DWLN_CASE(2) LVAE_DWLN_CASE(3) LVAE_DWLN_CASE(4)
    LVAE_DWLN_CASE(6)
    // C = 512 as a 2-CTA cluster of 4 chunks each: 168 -> ~100 registers per thread doubles the resident CTAs
    // (measured 25-40 % faster on the s16 / s32 / s64 layers; C = 384 was not faster split)
    case 8: return dispatch_k<4, 2>(k, x, dw_w, dw_b, ada, ada_stride, ada_off, ln_w, ln_b, y, p0, p1, p2, f16, B, H, W, st);
    case 10: return dispatch_k<5, 2>(k, x, dw_w, dw_b, ada, ada_stride, ada_off, ln_w, ln_b, y, p0, p1, p2, f16, B, H, W, st);
    case 12: return dispatch_k<6, 2>(k, x, dw_w, dw_b, ada, ada_stride, ada_off, ln_w, ln_b, y, p0, p1, p2, f16, B, H, W, st);
    default: set_error("dwconv channel count %d unsupported (need C/64 in {1,2,3,4,6,8,10,12})", C); return LVAE_E_UNSUPPORTED;
  }
#undef LVAE_DWLN_CASE
}

extern "C" int lvae_dwconv_ln_adaln(const float* x, const float* dw_w, const float* dw_b,
                                    const float* ada, int64_t ada_stride, int64_t ada_off,
                                    const float* ln_w, const float* ln_b,
                                    float* y, int B, int H, int W, int C, int k, void* stream) {
  LVAE_CHECK_ARG(y != nullptr);
  return dwln_dispatch(x, dw_w, dw_b, ada, ada_stride, ada_off, ln_w, ln_b, y, nullptr, nullptr, nullptr, 0, B, H, W, C, k, stream);
}

extern "C" int lvae_dwconv_ln_adaln_planes(const float* x, const float* dw_w, const float* dw_b,
                                           const float* ada, int64_t ada_stride, int64_t ada_off,
                                           const float* ln_w, const float* ln_b,
                                           void* y0, void* y1, void* y2, int plane_format,
                                           int B, int H, int W, int C, int k, void* stream) {
  LVAE_CHECK_ARG(y0 != nullptr && (y2 == nullptr || y1 != nullptr));
  LVAE_CHECK_ARG(plane_format == LVAE_PLANES_BF16 || plane_format == LVAE_PLANES_F16);
  return dwln_dispatch(x, dw_w, dw_b, ada, ada_stride, ada_off, ln_w, ln_b, nullptr, y0, y1, y2,
                       plane_format == LVAE_PLANES_F16, B, H, W, C, k, stream);
}
