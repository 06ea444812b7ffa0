// Backward of the depthwise-conv + LayerNorm + modulation stage of a ConvNeXt block (training step, SURVEY 8(a) a4
// under autograd; reference lvae/models/common.py:145-152, qresvae/model.py:163-182), as three HBM-bound kernels:
//
//   lvae_dwconv          c = dwconv_kxk(x) + bias           (recomputation of the conv output; and, with flip = 1 and
//                                                            `add`, the data gradient dx = conv(dc, flipped w) + add)
//   lvae_ln_mod_bwd      dc = LayerNorm'(c) . (da * g1), plus the gradients of the modulation (AdaLN shift / scale per
//                        image, or the affine LayerNorm's bias / weight)
//   lvae_dwconv_wgrad    dw[t, c] = sum_p dc[p, c] x[p + t, c],  db[c] = sum_p dc[p, c]
//
// All tensors NHWC fp32; the filter is packed [k*k, C] like the forward kernel's.  Tiles: 8 x 8 output pixels x 64
// channels (8 warps: warp = output column, lane = channel pair).  The two convolution kernels are persistent over the
// tiles of one 64-channel group; the (8+k-1)^2 halo of a tile is ONE TMA 4-D box load (cp.async.bulk.tensor over the
// (C, W, H, B) view, out-of-image coordinates zero-filled by the hardware = the conv's padding), double-buffered on two
// mbarriers so that tile t+1 streams in while tile t is computed.  Every input row of the halo is read from shared
// memory once per thread and feeds up to k output rows (k*k*8 packed FFMA2 per (8+k-1)*k LDS.64).
#include "common.cuh"
#include <cuda.h>
#include <mutex>

namespace lvae {

constexpr int DB_T = 8;        // tile edge
constexpr int DB_CH = 64;      // channels per CTA
constexpr int DB_THREADS = 256;

// ---- TMA plumbing: the halo tile is one 4-D box load over the (C, W, H, B) view of the NHWC tensor; coordinates outside
// the image are zero-filled by the hardware (= the conv's padding), so there is no fill loop and no index arithmetic
__device__ __forceinline__ uint32_t db_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void db_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void db_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void db_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void db_tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

// y = dwconv(x) [+ bias] [+ add];  FLIP: correlate with the spatially flipped filter (the transposed convolution).
// Persistent over the tiles of one 64-channel group, halo loads double-buffered (tile t+1 streams in while t is computed).
template <int K, bool FLIP>
__global__ void __launch_bounds__(DB_THREADS, 2) dwconv_kernel(const __grid_constant__ CUtensorMap x_map, const float* __restrict__ w,
                                                               const float* __restrict__ bias, const float* __restrict__ add,
                                                               float* __restrict__ y, int H, int W, int C, int tiles_x, int tiles_y,
                                                               int n_tiles) {
  constexpr int HT = DB_T + K - 1, PAD = (K - 1) / 2, TILE_FLOATS = HT * HT * DB_CH;
  extern __shared__ __align__(128) float xs[];                 // [2][HT][HT][64]
  __shared__ __align__(8) uint64_t bar[2];
  const int c0 = blockIdx.y * DB_CH;
  const int lane = threadIdx.x & 31, ox = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    db_mbar_init(db_smem_u32(&bar[0]), 1);
    db_mbar_init(db_smem_u32(&bar[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int tile, int buf) {                        // one thread
    const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
    const uint32_t bb = db_smem_u32(&bar[buf]);
    db_mbar_expect_tx(bb, (uint32_t)(TILE_FLOATS * 4));
    db_tma_load_4d(db_smem_u32(xs + buf * TILE_FLOATS), &x_map, bb, c0, tx * DB_T - PAD, ty * DB_T - PAD, b);
  };
  int tile = blockIdx.x;
  if (threadIdx.x == 0 && tile < n_tiles) issue(tile, 0);
  float2 wr[K * K];
#pragma unroll
  for (int t = 0; t < K * K; ++t)
    wr[t] = __ldg(reinterpret_cast<const float2*>(w + (int64_t)(FLIP ? K * K - 1 - t : t) * C + c0) + lane);
  const float2 b2 = bias ? __ldg(reinterpret_cast<const float2*>(bias + c0) + lane) : make_float2(0.f, 0.f);
  for (int it = 0; tile < n_tiles; tile += gridDim.x, ++it) {
    const int buf = it & 1;
    // buffer buf ^ 1 was last read in iteration it - 1, which ended with __syncthreads()
    if (threadIdx.x == 0 && tile + (int)gridDim.x < n_tiles) issue(tile + gridDim.x, buf ^ 1);
    const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
    float2 acc[DB_T];
#pragma unroll
    for (int i = 0; i < DB_T; ++i) acc[i] = b2;
    db_mbar_wait(db_smem_u32(&bar[buf]), (uint32_t)((it >> 1) & 1));
    const float2* xs2 = reinterpret_cast<const float2*>(xs + buf * TILE_FLOATS);
#pragma unroll
    for (int iy = 0; iy < HT; ++iy) {
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        const float2 v = xs2[(iy * HT + ox + kx) * (DB_CH / 2) + lane];
#pragma unroll
        for (int ky = 0; ky < K; ++ky) {
          const int oy = iy - ky;
          if (oy >= 0 && oy < DB_T) acc[oy] = fma2(v, wr[ky * K + kx], acc[oy]);
        }
      }
    }
    const int gx = tx * DB_T + ox;
    if (gx < W) {
      const int64_t o0 = (((int64_t)b * H + ty * DB_T) * W + gx) * C + c0 + 2 * lane;
      const int64_t rs = (int64_t)W * C;                       // one image row down
      const int rows = min(DB_T, H - ty * DB_T);
      if (add) {                                               // all residual loads in flight before the first is consumed
        float2 av[DB_T];
#pragma unroll
        for (int oy = 0; oy < DB_T; ++oy)
          av[oy] = oy < rows ? __ldg(reinterpret_cast<const float2*>(add + o0 + oy * rs)) : make_float2(0.f, 0.f);
#pragma unroll
        for (int oy = 0; oy < DB_T; ++oy) acc[oy] = add2(acc[oy], av[oy]);
      }
#pragma unroll
      for (int oy = 0; oy < DB_T; ++oy)
        if (oy < rows) *reinterpret_cast<float2*>(y + o0 + oy * rs) = acc[oy];
    }
    __syncthreads();
  }
}

// dw[t, c] += sum over the CTA's tiles of dc[p, c] * x[p + t, c];  db[c] += sum dc[p, c].  Same pipeline, two boxes per
// tile (the x halo and the 8 x 8 dc tile) on one mbarrier.
template <int K>
__global__ void __launch_bounds__(DB_THREADS, 1) dwconv_wgrad_kernel(const __grid_constant__ CUtensorMap x_map,
                                                                     const __grid_constant__ CUtensorMap dc_map,
                                                                     float* __restrict__ dw, float* __restrict__ db,
                                                                     int C, int tiles_x, int tiles_y, int n_tiles) {
  constexpr int HT = DB_T + K - 1, PAD = (K - 1) / 2;
  constexpr int X_FLOATS = HT * HT * DB_CH, D_FLOATS = DB_T * DB_T * DB_CH, BUF_FLOATS = X_FLOATS + D_FLOATS;
  extern __shared__ __align__(128) float smem[];              // [2][x halo | dc tile], then red[K*K + 1][64]
  __shared__ __align__(8) uint64_t bar[2];
  float* red = smem + 2 * BUF_FLOATS;
  const int c0 = blockIdx.y * DB_CH;
  const int lane = threadIdx.x & 31, ox = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    db_mbar_init(db_smem_u32(&bar[0]), 1);
    db_mbar_init(db_smem_u32(&bar[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < (K * K + 1) * DB_CH; i += DB_THREADS) red[i] = 0.f;
  __syncthreads();
  auto issue = [&](int tile, int buf) {
    const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
    const uint32_t bb = db_smem_u32(&bar[buf]);
    db_mbar_expect_tx(bb, (uint32_t)(BUF_FLOATS * 4));
    db_tma_load_4d(db_smem_u32(smem + buf * BUF_FLOATS), &x_map, bb, c0, tx * DB_T - PAD, ty * DB_T - PAD, b);
    db_tma_load_4d(db_smem_u32(smem + buf * BUF_FLOATS + X_FLOATS), &dc_map, bb, c0, tx * DB_T, ty * DB_T, b);
  };
  int tile = blockIdx.x;
  if (threadIdx.x == 0 && tile < n_tiles) issue(tile, 0);
  float2 acc[K * K];
#pragma unroll
  for (int t = 0; t < K * K; ++t) acc[t] = make_float2(0.f, 0.f);
  float2 bacc = make_float2(0.f, 0.f);
  for (int it = 0; tile < n_tiles; tile += gridDim.x, ++it) {
    const int buf = it & 1;
    if (threadIdx.x == 0 && tile + (int)gridDim.x < n_tiles) issue(tile + gridDim.x, buf ^ 1);
    db_mbar_wait(db_smem_u32(&bar[buf]), (uint32_t)((it >> 1) & 1));
    const float2* xs2 = reinterpret_cast<const float2*>(smem + buf * BUF_FLOATS);
    const float2* ds2 = reinterpret_cast<const float2*>(smem + buf * BUF_FLOATS + X_FLOATS);
    float2 d[DB_T];
#pragma unroll
    for (int oy = 0; oy < DB_T; ++oy) {
      d[oy] = ds2[(oy * DB_T + ox) * (DB_CH / 2) + lane];
      bacc = add2(bacc, d[oy]);
    }
#pragma unroll
    for (int iy = 0; iy < HT; ++iy) {
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        const float2 v = xs2[(iy * HT + ox + kx) * (DB_CH / 2) + lane];
#pragma unroll
        for (int ky = 0; ky < K; ++ky) {
          const int oy = iy - ky;
          if (oy >= 0 && oy < DB_T) acc[ky * K + kx] = fma2(d[oy], v, acc[ky * K + kx]);
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int t = 0; t < K * K; ++t) {
    atomicAdd(&red[t * DB_CH + 2 * lane], acc[t].x);
    atomicAdd(&red[t * DB_CH + 2 * lane + 1], acc[t].y);
  }
  atomicAdd(&red[K * K * DB_CH + 2 * lane], bacc.x);
  atomicAdd(&red[K * K * DB_CH + 2 * lane + 1], bacc.y);
  __syncthreads();
  for (int i = threadIdx.x; i < K * K * DB_CH; i += DB_THREADS)
    atomicAdd(dw + (int64_t)(i / DB_CH) * C + c0 + i % DB_CH, red[i]);
  if (threadIdx.x < DB_CH) atomicAdd(db + c0 + threadIdx.x, red[K * K * DB_CH + threadIdx.x]);
}

// LayerNorm (eps 1e-6, statistics over C) + modulation a = yhat * g1 + g0, backward.  One warp per pixel row, lane l owns
// channels {64 j + 2 l, 64 j + 2 l + 1}, j < NV = C / 64.  g1 = 1 + scale[b, c] (AdaLN) or ln_w[c] (affine).
// dmod rows (d g0 | d g1): row b for AdaLN ((dshift | dscale) of image b), row 0 for the affine parameters.
constexpr int LB_ROWS = 32;     // pixel rows per CTA
template <int NV>
__global__ void __launch_bounds__(256) ln_mod_bwd_kernel(const float* __restrict__ c, const float* __restrict__ da,
                                                         const float* __restrict__ ada, int64_t ada_stride, int64_t ada_off,
                                                         const float* __restrict__ ln_w, float* __restrict__ dc,
                                                         float* __restrict__ dmod, int HW) {
  constexpr int C = NV * 64;
  __shared__ float red[2 * C];
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 2 * C; i += 256) red[i] = 0.f;
  __syncthreads();
  float2 g1[NV], a0[NV], a1[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int ch = 64 * j + 2 * lane;
    if (ln_w) {
      g1[j] = __ldg(reinterpret_cast<const float2*>(ln_w + ch));
    } else {
      const float2 s = __ldg(reinterpret_cast<const float2*>(ada + b * ada_stride + ada_off + C + ch));
      g1[j] = make_float2(1.f + s.x, 1.f + s.y);
    }
    a0[j] = a1[j] = make_float2(0.f, 0.f);
  }
  const int r_end = min(HW, (int)(blockIdx.x + 1) * LB_ROWS);
  for (int r = blockIdx.x * LB_ROWS + warp; r < r_end; r += 8) {
    const int64_t o = ((int64_t)b * HW + r) * C;
    float2 cv[NV], dv[NV];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      cv[j] = __ldg(reinterpret_cast<const float2*>(c + o + 64 * j) + lane);
      dv[j] = __ldg(reinterpret_cast<const float2*>(da + o + 64 * j) + lane);
      s += cv[j].x + cv[j].y;
    }
    const float mean = warp_sum(s) * (1.f / C);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      cv[j].x -= mean; cv[j].y -= mean;
      q = fmaf(cv[j].x, cv[j].x, q); q = fmaf(cv[j].y, cv[j].y, q);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / C) + 1e-6f);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      cv[j].x *= rstd; cv[j].y *= rstd;                          // yhat
      a0[j].x += dv[j].x; a0[j].y += dv[j].y;
      a1[j].x = fmaf(dv[j].x, cv[j].x, a1[j].x); a1[j].y = fmaf(dv[j].y, cv[j].y, a1[j].y);
      dv[j].x *= g1[j].x; dv[j].y *= g1[j].y;                    // d yhat
      s1 += dv[j].x + dv[j].y;
      s2 = fmaf(dv[j].x, cv[j].x, s2); s2 = fmaf(dv[j].y, cv[j].y, s2);
    }
    s1 = warp_sum(s1) * (1.f / C);
    s2 = warp_sum(s2) * (1.f / C);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      float2 g;
      g.x = rstd * (dv[j].x - s1 - cv[j].x * s2);
      g.y = rstd * (dv[j].y - s1 - cv[j].y * s2);
      *(reinterpret_cast<float2*>(dc + o + 64 * j) + lane) = g;
    }
  }
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int ch = 64 * j + 2 * lane;
    atomicAdd(&red[ch], a0[j].x); atomicAdd(&red[ch + 1], a0[j].y);
    atomicAdd(&red[C + ch], a1[j].x); atomicAdd(&red[C + ch + 1], a1[j].y);
  }
  __syncthreads();
  float* out = dmod + (ln_w ? 0 : (int64_t)b * 2 * C);
  for (int i = threadIdx.x; i < 2 * C; i += 256) atomicAdd(out + i, red[i]);
}

typedef CUresult (*DbEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// NHWC fp32 viewed as a 4-D tensor (C, W, H, B); box = (64 channels, edge, edge, 1); out-of-image -> 0
static int db_make_map(CUtensorMap* map, const float* x, int B, int H, int W, int C, int edge) {
  static DbEncodeTiledFn enc = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      enc = reinterpret_cast<DbEncodeTiledFn>(ptr);
  });
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return LVAE_E_UNSUPPORTED; }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  cuuint32_t box[4] = {DB_CH, (cuuint32_t)edge, (cuuint32_t)edge, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (dwconv backward) failed: %d", (int)r); return LVAE_E_BADARG; }
  return 0;
}

static int db_sm_count() {
  static int n = 0;
  if (n == 0) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); }
  return n;
}

template <int K, bool FLIP>
static int dwconv_launch(const float* x, const float* w, const float* bias, const float* add, float* y,
                         int B, int H, int W, int C, cudaStream_t st) {
  constexpr int HT = DB_T + K - 1;
  const int tiles_x = (W + DB_T - 1) / DB_T, tiles_y = (H + DB_T - 1) / DB_T, n_tiles = tiles_x * tiles_y * B;
  const size_t smem = sizeof(float) * 2 * HT * HT * DB_CH;
  static bool once = false;
  if (!once) {
    LVAE_CUDA_CALL(cudaFuncSetAttribute(dwconv_kernel<K, FLIP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    once = true;
  }
  CUtensorMap map;
  int rc = db_make_map(&map, x, B, H, W, C, HT);
  if (rc) return rc;
  const int groups = C / DB_CH;
  int gx = (2 * db_sm_count() + groups - 1) / groups;     // two resident CTAs per SM, each looping over its share of the tiles
  if (gx > n_tiles) gx = n_tiles;
  dwconv_kernel<K, FLIP><<<dim3(gx, groups), DB_THREADS, smem, st>>>(map, w, bias, add, y, H, W, C, tiles_x, tiles_y, n_tiles);
  LVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

template <int K>
static int dwconv_wgrad_launch(const float* dc, const float* x, float* dw, float* db, int B, int H, int W, int C, cudaStream_t st) {
  constexpr int HT = DB_T + K - 1;
  const int tiles_x = (W + DB_T - 1) / DB_T, tiles_y = (H + DB_T - 1) / DB_T, n_tiles = tiles_x * tiles_y * B;
  const size_t smem = sizeof(float) * (2 * (HT * HT * DB_CH + DB_T * DB_T * DB_CH) + (K * K + 1) * DB_CH);
  static bool once = false;
  if (!once) {
    LVAE_CUDA_CALL(cudaFuncSetAttribute(dwconv_wgrad_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    once = true;
  }
  CUtensorMap xmap, dmap;
  int rc = db_make_map(&xmap, x, B, H, W, C, HT);
  if (rc) return rc;
  if ((rc = db_make_map(&dmap, dc, B, H, W, C, DB_T))) return rc;
  LVAE_CUDA_CALL(cudaMemsetAsync(dw, 0, sizeof(float) * K * K * C, st));
  LVAE_CUDA_CALL(cudaMemsetAsync(db, 0, sizeof(float) * C, st));
  const int groups = C / DB_CH;
  int gx = (db_sm_count() + groups - 1) / groups;         // one CTA per SM (146 KB of shared memory at k = 7)
  if (gx > n_tiles) gx = n_tiles;
  dwconv_wgrad_kernel<K><<<dim3(gx, groups), DB_THREADS, smem, st>>>(xmap, dmap, dw, db, C, tiles_x, tiles_y, n_tiles);
  LVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

}  // namespace lvae

using namespace lvae;

extern "C" int lvae_dwconv(const float* x, const float* dw_w, const float* bias, const float* add, float* y,
                           int B, int H, int W, int C, int k, int flip, void* stream) {
  LVAE_CHECK_ARG(x && dw_w && y && B > 0 && H > 0 && W > 0 && C > 0 && C % DB_CH == 0);
  LVAE_CHECK_ARG(k == 1 || k == 3 || k == 5 || k == 7);
  cudaStream_t st = (cudaStream_t)stream;
#define LVAE_DWC(K)                                                                                    \
  case K: return flip ? dwconv_launch<K, true>(x, dw_w, bias, add, y, B, H, W, C, st)                  \
                      : dwconv_launch<K, false>(x, dw_w, bias, add, y, B, H, W, C, st);
  switch (k) { LVAE_DWC(1) LVAE_DWC(3) LVAE_DWC(5) LVAE_DWC(7) }
#undef LVAE_DWC
  return LVAE_E_BADARG;
}

extern "C" int lvae_dwconv_wgrad(const float* dc, const float* x, float* dw, float* db,
                                 int B, int H, int W, int C, int k, void* stream) {
  LVAE_CHECK_ARG(dc && x && dw && db && B > 0 && H > 0 && W > 0 && C > 0 && C % DB_CH == 0);
  LVAE_CHECK_ARG(k == 1 || k == 3 || k == 5 || k == 7);
  cudaStream_t st = (cudaStream_t)stream;
  switch (k) {
    case 1: return dwconv_wgrad_launch<1>(dc, x, dw, db, B, H, W, C, st);
    case 3: return dwconv_wgrad_launch<3>(dc, x, dw, db, B, H, W, C, st);
    case 5: return dwconv_wgrad_launch<5>(dc, x, dw, db, B, H, W, C, st);
    case 7: return dwconv_wgrad_launch<7>(dc, x, dw, db, B, H, W, C, st);
  }
  return LVAE_E_BADARG;
}

extern "C" int lvae_ln_mod_bwd(const float* c, const float* da, const float* ada, int64_t ada_stride, int64_t ada_off,
                               const float* ln_w, float* dc, float* dmod, int B, int HW, int C, void* stream) {
  LVAE_CHECK_ARG(c && da && dc && dmod && (ada || ln_w) && B > 0 && HW > 0 && C > 0 && C % 64 == 0 && C <= 768);
  cudaStream_t st = (cudaStream_t)stream;
  LVAE_CUDA_CALL(cudaMemsetAsync(dmod, 0, sizeof(float) * 2 * C * (ln_w ? 1 : B), st));
  const dim3 grid((HW + LB_ROWS - 1) / LB_ROWS, B);
#define LVAE_LNB(NV)                                                                                             \
  case NV: ln_mod_bwd_kernel<NV><<<grid, 256, 0, st>>>(c, da, ada, ada_stride, ada_off, ln_w, dc, dmod, HW); break;
  switch (C / 64) {
    LVAE_LNB(1) LVAE_LNB(2) LVAE_LNB(3) LVAE_LNB(4) LVAE_LNB(5) LVAE_LNB(6) LVAE_LNB(7) LVAE_LNB(8) LVAE_LNB(9)
    LVAE_LNB(10) LVAE_LNB(11) LVAE_LNB(12)
    default: return LVAE_E_BADARG;
  }
#undef LVAE_LNB
  LVAE_CUDA_LAUNCH_CHECK();
  return 0;
}
