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
// channels per CTA (8 warps: warp = output column, lane = channel pair), the (8+k-1)^2 halo staged in shared memory
// with zero fill = the conv's padding.  Every input row of the halo is read from shared memory once per thread and feeds
// up to k output rows (k*k*8 packed FFMA2 per (8+k-1)*k LDS.64).
#include "common.cuh"
#include <cuda.h>

namespace lvae {

constexpr int DB_T = 8;        // tile edge
constexpr int DB_CH = 64;      // channels per CTA
constexpr int DB_THREADS = 256;

template <int K>
__device__ __forceinline__ void db_load_halo(float* xs, const float* __restrict__ x, int b, int ty, int tx, int c0,
                                             int H, int W, int C) {
  constexpr int HT = DB_T + K - 1, PAD = (K - 1) / 2;
  for (int i = threadIdx.x; i < HT * HT * (DB_CH / 4); i += DB_THREADS) {
    const int pix = i / (DB_CH / 4), c4 = i % (DB_CH / 4);
    const int gy = ty * DB_T + pix / HT - PAD, gx = tx * DB_T + pix % HT - PAD;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gy >= 0 && gy < H && gx >= 0 && gx < W)
      v = __ldg(reinterpret_cast<const float4*>(x + (((int64_t)b * H + gy) * W + gx) * C + c0 + c4 * 4));
    reinterpret_cast<float4*>(xs)[i] = v;
  }
}

// y = dwconv(x) [+ bias] [+ add];  FLIP: correlate with the spatially flipped filter (the transposed convolution)
template <int K, bool FLIP>
__global__ void __launch_bounds__(DB_THREADS, 2) dwconv_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ bias, const float* __restrict__ add,
                                                            float* __restrict__ y, int H, int W, int C, int tiles_x, int tiles_y) {
  constexpr int HT = DB_T + K - 1;
  extern __shared__ __align__(16) float xs[];
  const int tile = blockIdx.x, c0 = blockIdx.y * DB_CH;
  const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
  const int lane = threadIdx.x & 31, ox = threadIdx.x >> 5;
  db_load_halo<K>(xs, x, b, ty, tx, c0, H, W, C);
  float2 wr[K * K];
#pragma unroll
  for (int t = 0; t < K * K; ++t)
    wr[t] = __ldg(reinterpret_cast<const float2*>(w + (int64_t)(FLIP ? K * K - 1 - t : t) * C + c0) + lane);
  float2 acc[DB_T];
  const float2 b2 = bias ? __ldg(reinterpret_cast<const float2*>(bias + c0) + lane) : make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < DB_T; ++i) acc[i] = b2;
  __syncthreads();
  const float2* xs2 = reinterpret_cast<const float2*>(xs);
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
#pragma unroll
    for (int oy = 0; oy < DB_T; ++oy) {
      const int gy = ty * DB_T + oy;
      if (gy < H) {
        const int64_t o = (((int64_t)b * H + gy) * W + gx) * C + c0 + 2 * lane;
        float2 r = acc[oy];
        if (add) r = add2(r, __ldg(reinterpret_cast<const float2*>(add + o)));
        *reinterpret_cast<float2*>(y + o) = r;
      }
    }
  }
}

// dw[t, c] += sum over the CTA's tiles of dc[p, c] * x[p + t, c];  db[c] += sum dc[p, c]
template <int K>
__global__ void __launch_bounds__(DB_THREADS, 2) dwconv_wgrad_kernel(const float* __restrict__ dc, const float* __restrict__ x,
                                                                  float* __restrict__ dw, float* __restrict__ db,
                                                                  int H, int W, int C, int tiles_x, int tiles_y, int n_tiles) {
  constexpr int HT = DB_T + K - 1;
  extern __shared__ __align__(16) float smem[];
  float* xs = smem;                               // [HT][HT][64]
  float* ds = smem + HT * HT * DB_CH;             // [8][8][64]
  float* red = ds + DB_T * DB_T * DB_CH;          // [K*K + 1][64]
  const int c0 = blockIdx.y * DB_CH;
  const int lane = threadIdx.x & 31, ox = threadIdx.x >> 5;
  float2 acc[K * K];
#pragma unroll
  for (int t = 0; t < K * K; ++t) acc[t] = make_float2(0.f, 0.f);
  float2 bacc = make_float2(0.f, 0.f);
  for (int i = threadIdx.x; i < (K * K + 1) * DB_CH; i += DB_THREADS) red[i] = 0.f;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
    __syncthreads();                              // previous tile's reads are done
    db_load_halo<K>(xs, x, b, ty, tx, c0, H, W, C);
    for (int i = threadIdx.x; i < DB_T * DB_T * (DB_CH / 4); i += DB_THREADS) {
      const int pix = i / (DB_CH / 4), c4 = i % (DB_CH / 4);
      const int gy = ty * DB_T + pix / DB_T, gx = tx * DB_T + pix % DB_T;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gy < H && gx < W)
        v = __ldg(reinterpret_cast<const float4*>(dc + (((int64_t)b * H + gy) * W + gx) * C + c0 + c4 * 4));
      reinterpret_cast<float4*>(ds)[i] = v;
    }
    __syncthreads();
    const float2* xs2 = reinterpret_cast<const float2*>(xs);
    const float2* ds2 = reinterpret_cast<const float2*>(ds);
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
  }
  __syncthreads();
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

template <int K, bool FLIP>
static int dwconv_launch(const float* x, const float* w, const float* bias, const float* add, float* y,
                         int B, int H, int W, int C, cudaStream_t st) {
  constexpr int HT = DB_T + K - 1;
  const int tiles_x = (W + DB_T - 1) / DB_T, tiles_y = (H + DB_T - 1) / DB_T;
  const size_t smem = sizeof(float) * HT * HT * DB_CH;
  static bool once = false;
  if (!once) {
    LVAE_CUDA_CALL(cudaFuncSetAttribute(dwconv_kernel<K, FLIP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    once = true;
  }
  dwconv_kernel<K, FLIP><<<dim3(tiles_x * tiles_y * B, C / DB_CH), DB_THREADS, smem, st>>>(x, w, bias, add, y, H, W, C, tiles_x, tiles_y);
  LVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

template <int K>
static int dwconv_wgrad_launch(const float* dc, const float* x, float* dw, float* db, int B, int H, int W, int C, cudaStream_t st) {
  constexpr int HT = DB_T + K - 1;
  const int tiles_x = (W + DB_T - 1) / DB_T, tiles_y = (H + DB_T - 1) / DB_T, n_tiles = tiles_x * tiles_y * B;
  const size_t smem = sizeof(float) * (HT * HT * DB_CH + DB_T * DB_T * DB_CH + (K * K + 1) * DB_CH);
  static bool once = false;
  if (!once) {
    LVAE_CUDA_CALL(cudaFuncSetAttribute(dwconv_wgrad_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    once = true;
  }
  LVAE_CUDA_CALL(cudaMemsetAsync(dw, 0, sizeof(float) * K * K * C, st));
  LVAE_CUDA_CALL(cudaMemsetAsync(db, 0, sizeof(float) * C, st));
  const int groups = C / DB_CH;
  int gx = (2 * 148 + groups - 1) / groups;       // about two CTAs per SM in total, each looping over its share of the tiles
  if (gx > n_tiles) gx = n_tiles;
  dwconv_wgrad_kernel<K><<<dim3(gx, groups), DB_THREADS, smem, st>>>(dc, x, dw, db, H, W, C, tiles_x, tiles_y, n_tiles);
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
