// Depthwise kxk conv + LayerNorm(C, eps 1e-6) + AdaLN (or affine LN), NHWC fp32 in, fp32 and/or bf16 planes out.
// Reference: ConvNeXtBlockAdaLN.forward, lvae/models/common.py:145-152 (conv_dw -> permute -> norm ->
// x*(1+scale)+shift); qresvae MyConvNeXtBlock (affine LayerNorm, no AdaLN) qresvae/model.py:163-182.
//
// HBM-bound stage (algorithmic bytes per position: read 4C, write 4C fp32 or 2C per bf16 plane).  One CTA of
// 8 warps owns an 8 x 8 tile of output pixels of one image, all C channels:
//   * channels are processed in chunks of 64; the (8+k-1)^2 halo of a chunk is staged in shared memory with
//     16-byte cp.async (zero-filled outside the image), double-buffered so chunk j+1 streams in while chunk j
//     is convolved -- every input element is read from global memory once per tile, fully coalesced;
//   * warp w convolves output row w: lane l owns channels {64 j + 2 l, 64 j + 2 l + 1}, a sliding window of the
//     shared-memory row feeds the 8 pixels, results stay in registers (8 pixels x 2 channels x C/64 chunks);
//   * LayerNorm is then warp-local: two-pass (mean, centred second moment) warp-shuffle reductions in fp32,
//     followed by the modulation and the split into bf16 planes for the tensor-core GEMM that consumes it.
#include "common.cuh"
#include <cuda_bf16.h>

namespace lvae {

constexpr int DW_T = 8;                 // output tile edge
constexpr int DW_CH = 64;               // channels per chunk
constexpr int DW_NBUF = 1;              // halo buffers per CTA: 1 = rely on co-resident CTAs for overlap (more CTAs per SM)

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;        // src-size 0: the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// packed fp32x2 FMA (sm_100): both lanes are IEEE fma, i.e. bit-identical to two fmaf() at half the issue slots
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d)
      : "l"(*reinterpret_cast<const uint64_t*>(&a)), "l"(*reinterpret_cast<const uint64_t*>(&b)),
        "l"(*reinterpret_cast<const uint64_t*>(&c)));
  return *reinterpret_cast<float2*>(&d);
}
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NJ, int KS>
__global__ void __launch_bounds__(256) dwln_kernel(
    const float* __restrict__ x, const float* __restrict__ dw_w, const float* __restrict__ dw_b,
    const float* __restrict__ ada, int64_t ada_stride, int64_t ada_off,
    const float* __restrict__ ln_w, const float* __restrict__ ln_b,
    float* __restrict__ y, __nv_bfloat16* __restrict__ y0, __nv_bfloat16* __restrict__ y1, __nv_bfloat16* __restrict__ y2,
    int H, int W, int tiles_x, int tiles_y) {
  constexpr int C = NJ * DW_CH, PAD = (KS - 1) / 2, HT = DW_T + KS - 1;     // halo tile edge
  constexpr int CHUNK_FLOATS = HT * HT * DW_CH;
  extern __shared__ __align__(16) float dw_smem[];                           // [2][HT][HT][64]
  const int tid = threadIdx.x, lane = tid & 31, wrow = tid >> 5;
  int t = blockIdx.x;
  const int tx = t % tiles_x; t /= tiles_x;
  const int ty = t % tiles_y; const int b = t / tiles_y;
  const int h0 = ty * DW_T, w0 = tx * DW_T;
  const float* xb = x + (int64_t)b * H * W * C;

  auto load_chunk = [&](int j, int buf) {
    // HT*HT pixels x 16 float4 per pixel
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(dw_smem + buf * CHUNK_FLOATS);
    for (int i = tid; i < HT * HT * 16; i += 256) {
      const int q = i & 15, pix = i >> 4;
      const int px = pix % HT, py = pix / HT;
      const int hh = h0 + py - PAD, ww = w0 + px - PAD;
      const bool ok = hh >= 0 && hh < H && ww >= 0 && ww < W;
      const float* src = ok ? xb + ((int64_t)hh * W + ww) * C + j * DW_CH + q * 4 : xb;
      cp_async16(sbase + (uint32_t)(pix * DW_CH + q * 4) * 4u, src, ok);
    }
    cp_async_commit();
  };

  float2 res[NJ][DW_T];
  if (DW_NBUF == 2) load_chunk(0, 0);
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    if (DW_NBUF == 2) {
      if (j + 1 < NJ) { load_chunk(j + 1, (j + 1) & 1); cp_async_wait<1>(); }
      else cp_async_wait<0>();
    } else {
      load_chunk(j, 0);
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* tile = dw_smem + (DW_NBUF == 2 ? (j & 1) : 0) * CHUNK_FLOATS;
    const int c = j * DW_CH + lane * 2;
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
  if (h >= H) return;                      // warp-uniform
  float mean[DW_T], rstd[DW_T];
#pragma unroll
  for (int s = 0; s < DW_T; ++s) {
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) sum += res[j][s].x + res[j][s].y;
    mean[s] = warp_sum(sum) * (1.0f / C);
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const float dx = res[j][s].x - mean[s], dy = res[j][s].y - mean[s];
      sq = fmaf(dx, dx, sq); sq = fmaf(dy, dy, sq);
    }
    rstd[s] = 1.0f / sqrtf(warp_sum(sq) * (1.0f / C) + 1e-6f);
  }
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int c = j * DW_CH + lane * 2;
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
        // A operand of the tensor-core fc1 GEMM: p0 = rn_bf16(v), p1 = rn_bf16(v - p0), p2 = rn_bf16(v - p0 - p1)
        const __nv_bfloat162 hv = __floats2bfloat162_rn(v.x, v.y);
        *reinterpret_cast<__nv_bfloat162*>(y0 + o) = hv;
        if (y1 != nullptr) {
          const float2 hf = __bfloat1622float2(hv);
          const float2 r1 = make_float2(__fsub_rn(v.x, hf.x), __fsub_rn(v.y, hf.y));
          const __nv_bfloat162 mv = __floats2bfloat162_rn(r1.x, r1.y);
          *reinterpret_cast<__nv_bfloat162*>(y1 + o) = mv;
          if (y2 != nullptr) {
            const float2 mf = __bfloat1622float2(mv);
            *reinterpret_cast<__nv_bfloat162*>(y2 + o) = __floats2bfloat162_rn(__fsub_rn(r1.x, mf.x), __fsub_rn(r1.y, mf.y));
          }
        }
      }
    }
  }
}

template <int NJ, int KS>
static int launch_dwln(const float* x, const float* dw_w, const float* dw_b, const float* ada,
                       int64_t ada_stride, int64_t ada_off, const float* ln_w, const float* ln_b,
                       float* y, __nv_bfloat16* y0, __nv_bfloat16* y1, __nv_bfloat16* y2,
                       int B, int H, int W, cudaStream_t stream) {
  constexpr int HT = DW_T + KS - 1;
  constexpr int smem = DW_NBUF * HT * HT * DW_CH * 4;
  static bool configured = false;
  if (!configured) {
    LVAE_CUDA_CALL(cudaFuncSetAttribute(dwln_kernel<NJ, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const int tiles_x = (W + DW_T - 1) / DW_T, tiles_y = (H + DW_T - 1) / DW_T;
  const int64_t blocks = (int64_t)B * tiles_x * tiles_y;
  dwln_kernel<NJ, KS><<<(unsigned)blocks, 256, smem, stream>>>(
      x, dw_w, dw_b, ada, ada_stride, ada_off, ln_w, ln_b, y, y0, y1, y2, H, W, tiles_x, tiles_y);
  LVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

template <int NJ>
static int dispatch_k(int k, const float* x, const float* dw_w, const float* dw_b, const float* ada,
                      int64_t ada_stride, int64_t ada_off, const float* ln_w, const float* ln_b,
                      float* y, __nv_bfloat16* y0, __nv_bfloat16* y1, __nv_bfloat16* y2,
                      int B, int H, int W, cudaStream_t stream) {
  switch (k) {
    case 1: return launch_dwln<NJ, 1>(x, dw_w, dw_b, ada, ada_stride, ada_off, ln_w, ln_b, y, y0, y1, y2, B, H, W, stream);
    case 3: return launch_dwln<NJ, 3>(x, dw_w, dw_b, ada, ada_stride, ada_off, ln_w, ln_b, y, y0, y1, y2, B, H, W, stream);
    case 5: return launch_dwln<NJ, 5>(x, dw_w, dw_b, ada, ada_stride, ada_off, ln_w, ln_b, y, y0, y1, y2, B, H, W, stream);
    case 7: return launch_dwln<NJ, 7>(x, dw_w, dw_b, ada, ada_stride, ada_off, ln_w, ln_b, y, y0, y1, y2, B, H, W, stream);
    default: set_error("dwconv kernel size %d unsupported", k); return LVAE_E_UNSUPPORTED;
  }
}

}  // namespace lvae

static int dwln_dispatch(const float* x, const float* dw_w, const float* dw_b,
                         const float* ada, int64_t ada_stride, int64_t ada_off,
                         const float* ln_w, const float* ln_b, float* y, void* y0, void* y1, void* y2,
                         int B, int H, int W, int C, int k, void* stream) {
  using namespace lvae;
  LVAE_CHECK_ARG(x && dw_w && dw_b && (y || y0) && (ada || ln_w));
  LVAE_CHECK_ARG(B > 0 && H > 0 && W > 0 && C > 0 && C % 64 == 0);
  LVAE_CHECK_ARG(ln_w != nullptr || (ada_off % 2 == 0 && ada_stride % 2 == 0));      // float2 loads of shift / scale
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* p0 = (__nv_bfloat16*)y0; __nv_bfloat16* p1 = (__nv_bfloat16*)y1; __nv_bfloat16* p2 = (__nv_bfloat16*)y2;
#define LVAE_DWLN_CASE(nj) case nj: return dispatch_k<nj>(k, x, dw_w, dw_b, ada, ada_stride, ada_off, ln_w, ln_b, y, p0, p1, p2, B, H, W, st);
  switch (C / 64) {
    LVAE_DWLN_CASE(1) LVAE_DWLN_CASE(2) LVAE_DWLN_CASE(3) LVAE_DWLN_CASE(4)
    LVAE_DWLN_CASE(6) LVAE_DWLN_CASE(8)
    default: set_error("dwconv channel count %d unsupported (need C/64 in {1,2,3,4,6,8})", C); return LVAE_E_UNSUPPORTED;
  }
#undef LVAE_DWLN_CASE
}

extern "C" int lvae_dwconv_ln_adaln(const float* x, const float* dw_w, const float* dw_b,
                                    const float* ada, int64_t ada_stride, int64_t ada_off,
                                    const float* ln_w, const float* ln_b,
                                    float* y, int B, int H, int W, int C, int k, void* stream) {
  LVAE_CHECK_ARG(y != nullptr);
  return dwln_dispatch(x, dw_w, dw_b, ada, ada_stride, ada_off, ln_w, ln_b, y, nullptr, nullptr, nullptr, B, H, W, C, k, stream);
}

extern "C" int lvae_dwconv_ln_adaln_planes(const float* x, const float* dw_w, const float* dw_b,
                                           const float* ada, int64_t ada_stride, int64_t ada_off,
                                           const float* ln_w, const float* ln_b,
                                           void* y0, void* y1, void* y2, int B, int H, int W, int C, int k, void* stream) {
  LVAE_CHECK_ARG(y0 != nullptr && (y2 == nullptr || y1 != nullptr));
  return dwln_dispatch(x, dw_w, dw_b, ada, ada_stride, ada_off, ln_w, ln_b, nullptr, y0, y1, y2, B, H, W, C, k, stream);
}
