// Depthwise kxk conv + LayerNorm(C, eps 1e-6) + AdaLN (or affine LN), NHWC fp32.
// Reference: ConvNeXtBlockAdaLN.forward, lvae/models/common.py:145-152 (conv_dw -> permute -> norm ->
// x*(1+scale)+shift); qresvae MyConvNeXtBlock (affine LayerNorm, no AdaLN) qresvae/model.py:163-182.
//
// HBM-bound stage: reads x once (neighbour re-reads hit L1/L2), writes the MLP's A operand once.
// One warp owns a strip of S consecutive output pixels of one image row; lane l owns channels
// {64 j + 2 l, 64 j + 2 l + 1}, j < C/64, so every global access of a warp is one contiguous 256 B
// line.  The strip re-uses each loaded input pixel for up to k outputs (sliding window in registers).
// LayerNorm is a two-pass (mean, then centred second moment) warp-shuffle reduction in fp32.
#include "common.cuh"
#include <cuda_bf16.h>

namespace lvae {

template <int NJ, int KS, int S>
__global__ void __launch_bounds__(256) dwln_kernel(
    const float* __restrict__ x, const float* __restrict__ dw_w, const float* __restrict__ dw_b,
    const float* __restrict__ ada, int64_t ada_stride, int64_t ada_off,
    const float* __restrict__ ln_w, const float* __restrict__ ln_b,
    float* __restrict__ y, __nv_bfloat16* __restrict__ y_hi, __nv_bfloat16* __restrict__ y_lo, __nv_bfloat16* __restrict__ y_l2,
    int B, int H, int W, int strips_per_row, int64_t total_strips) {
  constexpr int C = NJ * 64, PAD = (KS - 1) / 2, NX = S + KS - 1;
  const int lane = threadIdx.x & 31;
  const int64_t strip = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (strip >= total_strips) return;
  const int sw = (int)(strip % strips_per_row);
  const int64_t row = strip / strips_per_row;        // b*H + h
  const int h = (int)(row % H); const int b = (int)(row / H);
  const int w0 = sw * S;

  float2 res[NJ][S];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int c = j * 64 + lane * 2;
    const float2 bias = __ldg(reinterpret_cast<const float2*>(dw_b + c));
    float2 acc[S];
#pragma unroll
    for (int s = 0; s < S; ++s) acc[s] = bias;
#pragma unroll
    for (int ky = 0; ky < KS; ++ky) {
      const int hh = h + ky - PAD;
      if (hh < 0 || hh >= H) continue;     // warp-uniform
      const float* xrow = x + (((int64_t)b * H + hh) * W) * C + c;
      float2 xv[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        const int ww = w0 + i - PAD;
        xv[i] = (ww >= 0 && ww < W) ? __ldg(reinterpret_cast<const float2*>(xrow + (int64_t)ww * C))
                                    : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int kx = 0; kx < KS; ++kx) {
        const float2 wv = __ldg(reinterpret_cast<const float2*>(dw_w + (ky * KS + kx) * C + c));
#pragma unroll
        for (int s = 0; s < S; ++s) {
          acc[s].x = fmaf(xv[s + kx].x, wv.x, acc[s].x);
          acc[s].y = fmaf(xv[s + kx].y, wv.y, acc[s].y);
        }
      }
    }
#pragma unroll
    for (int s = 0; s < S; ++s) res[j][s] = acc[s];
  }

  // LayerNorm + modulation per output pixel
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const int w = w0 + s;
    if (w >= W) break;                      // warp-uniform
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) sum += res[j][s].x + res[j][s].y;
    const float mean = warp_sum(sum) * (1.0f / C);
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const float dx = res[j][s].x - mean, dy = res[j][s].y - mean;
      sq = fmaf(dx, dx, sq); sq = fmaf(dy, dy, sq);
    }
    const float var = warp_sum(sq) * (1.0f / C);
    const float rstd = 1.0f / sqrtf(var + 1e-6f);
    const int64_t yoff = (((int64_t)b * H + h) * W + w) * C;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int c = j * 64 + lane * 2;
      float2 v;
      v.x = __fmul_rn(res[j][s].x - mean, rstd);
      v.y = __fmul_rn(res[j][s].y - mean, rstd);
      if (ln_w != nullptr) {
        const float2 g = __ldg(reinterpret_cast<const float2*>(ln_w + c));
        const float2 be = __ldg(reinterpret_cast<const float2*>(ln_b + c));
        v.x = __fadd_rn(__fmul_rn(v.x, g.x), be.x);
        v.y = __fadd_rn(__fmul_rn(v.y, g.y), be.y);
      } else {
        const float* e = ada + (int64_t)b * ada_stride + ada_off + c;
        const float2 shift = __ldg(reinterpret_cast<const float2*>(e));
        const float2 scale = __ldg(reinterpret_cast<const float2*>(e + C));
        v.x = __fadd_rn(__fmul_rn(v.x, __fadd_rn(1.0f, scale.x)), shift.x);
        v.y = __fadd_rn(__fmul_rn(v.y, __fadd_rn(1.0f, scale.y)), shift.y);
      }
      if (y != nullptr) *reinterpret_cast<float2*>(y + yoff + c) = v;
      if (y_hi != nullptr) {
        // A operand of the tensor-core fc1 GEMM: hi = rn_bf16(v), lo = rn_bf16(v - hi)
        const __nv_bfloat162 hv = __floats2bfloat162_rn(v.x, v.y);
        *reinterpret_cast<__nv_bfloat162*>(y_hi + yoff + c) = hv;
        if (y_lo != nullptr) {
          const float2 hf = __bfloat1622float2(hv);
          const float2 r1 = make_float2(__fsub_rn(v.x, hf.x), __fsub_rn(v.y, hf.y));
          const __nv_bfloat162 mv = __floats2bfloat162_rn(r1.x, r1.y);
          *reinterpret_cast<__nv_bfloat162*>(y_lo + yoff + c) = mv;
          if (y_l2 != nullptr) {
            const float2 mf = __bfloat1622float2(mv);
            *reinterpret_cast<__nv_bfloat162*>(y_l2 + yoff + c) = __floats2bfloat162_rn(__fsub_rn(r1.x, mf.x), __fsub_rn(r1.y, mf.y));
          }
        }
      }
    }
  }
}

template <int NJ, int KS>
static int launch_dwln(const float* x, const float* dw_w, const float* dw_b, const float* ada,
                       int64_t ada_stride, int64_t ada_off, const float* ln_w, const float* ln_b,
                       float* y, __nv_bfloat16* y_hi, __nv_bfloat16* y_lo, __nv_bfloat16* y_l2, int B, int H, int W, cudaStream_t stream) {
  constexpr int S = (NJ >= 6) ? 4 : 4;
  const int spr = (W + S - 1) / S;
  const int64_t total = (int64_t)B * H * spr;
  const int warps = 8;
  const int64_t blocks = (total + warps - 1) / warps;
  dwln_kernel<NJ, KS, S><<<(unsigned)blocks, warps * 32, 0, stream>>>(
      x, dw_w, dw_b, ada, ada_stride, ada_off, ln_w, ln_b, y, y_hi, y_lo, y_l2, B, H, W, spr, total);
  LVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

template <int NJ>
static int dispatch_k(int k, const float* x, const float* dw_w, const float* dw_b, const float* ada,
                      int64_t ada_stride, int64_t ada_off, const float* ln_w, const float* ln_b,
                      float* y, __nv_bfloat16* y_hi, __nv_bfloat16* y_lo, __nv_bfloat16* y_l2, int B, int H, int W, cudaStream_t stream) {
  switch (k) {
    case 1: return launch_dwln<NJ, 1>(x, dw_w, dw_b, ada, ada_stride, ada_off, ln_w, ln_b, y, y_hi, y_lo, y_l2, B, H, W, stream);
    case 3: return launch_dwln<NJ, 3>(x, dw_w, dw_b, ada, ada_stride, ada_off, ln_w, ln_b, y, y_hi, y_lo, y_l2, B, H, W, stream);
    case 5: return launch_dwln<NJ, 5>(x, dw_w, dw_b, ada, ada_stride, ada_off, ln_w, ln_b, y, y_hi, y_lo, y_l2, B, H, W, stream);
    case 7: return launch_dwln<NJ, 7>(x, dw_w, dw_b, ada, ada_stride, ada_off, ln_w, ln_b, y, y_hi, y_lo, y_l2, B, H, W, stream);
    default: set_error("dwconv kernel size %d unsupported", k); return LVAE_E_UNSUPPORTED;
  }
}

}  // namespace lvae

static int dwln_dispatch(const float* x, const float* dw_w, const float* dw_b,
                         const float* ada, int64_t ada_stride, int64_t ada_off,
                         const float* ln_w, const float* ln_b, float* y, void* y_hi, void* y_lo, void* y_l2,
                         int B, int H, int W, int C, int k, void* stream) {
  using namespace lvae;
  LVAE_CHECK_ARG(x && dw_w && dw_b && (y || y_hi) && (ada || ln_w));
  LVAE_CHECK_ARG(B > 0 && H > 0 && W > 0 && C > 0 && C % 64 == 0);
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* h = (__nv_bfloat16*)y_hi; __nv_bfloat16* l = (__nv_bfloat16*)y_lo; __nv_bfloat16* l2 = (__nv_bfloat16*)y_l2;
#define LVAE_DWLN_CASE(nj) case nj: return dispatch_k<nj>(k, x, dw_w, dw_b, ada, ada_stride, ada_off, ln_w, ln_b, y, h, l, l2, B, H, W, st);
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
