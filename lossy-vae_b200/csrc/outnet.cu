// GaussianNLLOutputNet (qres34m_lossless: lvae/models/qresvae/model.py:16-94): the arithmetic AFTER the two
// patch_upsample heads conv_mean / conv_scale (those are lvae_gemm launches with LVAE_EPI_SHUFFLE_NCHW).  All tensors
// here are NCHW [B, 3, H, W] fp32, i.e. flat [B, chw].
//
//   forward_loss (model.py:24-40):   x = (im - 0.5) * 2;  ls = softplus(p_logscale + 16) - 16;
//       nll = -mean_chw gaussian_log_prob_mass(p_mean, exp(ls), x, bin = 1/127.5, prob_clamp = 1e-6)
//       (entropy_coding.py:17-49: mass = cdf(x + bin/2) - cdf(x - bin/2); log(max(mass, 1e-8)) where mass > 1e-6,
//        else Normal.log_prob(x) + log(bin)), same op order as td.Normal on the CPU (latent.cu's training branch)
//   _preapre_codec (model.py:68-79): pm = (round(p_mean * 127.5 + 127.5) / 127.5 - 1) / bin;  plogv = p_logscale - log(bin)
//   compress (model.py:81-86):       idx = build_indexes(exp(plogv)) on the 128-entry table; sym = round(x / bin - pm)
//   decompress (model.py:88-94):     x_hat = (sym + pm) * bin
// Per-image sums are deterministic per-block partials in the layout of lvae_image_distortion (lvae_image_num_partials).
#include "common.cuh"

namespace lvae {

constexpr int ON_T = 256, ON_E = 8;                  // = DT, DE of misc.cu: the same partial-sum layout
constexpr float ON_BIN = (float)(1.0 / 127.5);
constexpr float ON_HALF_BIN = (float)(0.5 * (1.0 / 127.5));      // python: 0.5 * bin_size, rounded once to fp32
constexpr float ON_LOG_BIN = -4.848116339763911f;                 // math.log(1 / 127.5)

__device__ __forceinline__ float on_softplus(float x) { return x > 20.0f ? x : log1pf(expf(x)); }   // beta 1, threshold 20
__device__ __forceinline__ float on_cdf(float v, float mean, float rcp) {
  // td.Normal(mean, scale).cdf(v) = 0.5 * (1 + erf((v - mean) * scale.reciprocal() / sqrt(2)))
  return __fmul_rn(0.5f, __fadd_rn(1.0f, erf_torch_cpu(__fdiv_rn(__fmul_rn(__fsub_rn(v, mean), rcp), 1.4142135623730951f))));
}

__global__ void __launch_bounds__(ON_T) nll_output_kernel(const float* __restrict__ p_mean, const float* __restrict__ p_ls,
                                                          const float* __restrict__ im, float* __restrict__ partial,
                                                          int chw, int nparts) {
  __shared__ float red[ON_T / 32];
  const int b = blockIdx.y;
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < ON_E; ++e) {
    const int i = (blockIdx.x * ON_E + e) * ON_T + threadIdx.x;
    if (i < chw) {
      const int64_t o = (int64_t)b * chw + i;
      const float x = __fmul_rn(__fadd_rn(im[o], -0.5f), 2.0f);                  // preprocess_target
      const float mean = p_mean[o];
      const float ls = __fsub_rn(on_softplus(__fadd_rn(p_ls[o], 16.0f)), 16.0f);
      const float scale = expf(ls);
      const float rcp = __frcp_rn(scale);
      const float mass = __fsub_rn(on_cdf(__fadd_rn(x, ON_HALF_BIN), mean, rcp), on_cdf(__fsub_rn(x, ON_HALF_BIN), mean, rcp));
      float lp;
      if (mass > 1e-6f) {
        lp = logf(fmaxf(mass, 1e-8f));
      } else {
        const float d = __fsub_rn(x, mean);
        const float var = __fmul_rn(scale, scale);
        lp = __fsub_rn(__fsub_rn(__fdiv_rn(-__fmul_rn(d, d), __fmul_rn(2.0f, var)), logf(scale)), 0.9189385332046727f);
        lp = __fadd_rn(lp, ON_LOG_BIN);
      }
      s += -lp;
    }
  }
  s = warp_sum(s);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) red[wid] = s;
  __syncthreads();
  if (wid == 0) {
    float a = lane < ON_T / 32 ? red[lane] : 0.f;
    a = warp_sum(a);
    if (lane == 0) partial[(int64_t)b * nparts + blockIdx.x] = a;
  }
}

// codec side: pm (in bins), table index of exp(plogv) and -- when im is given -- the residual symbol
__global__ void __launch_bounds__(256) outnet_codec_kernel(const float* __restrict__ p_mean, const float* __restrict__ p_ls,
                                                           const float* __restrict__ im, const float* __restrict__ table, int n_scales,
                                                           float* __restrict__ pm_out, int32_t* __restrict__ idx, int32_t* __restrict__ sym,
                                                           int64_t total) {
  __shared__ float stab[128];
  for (int i = threadIdx.x; i < n_scales && i < 128; i += blockDim.x) stab[i] = table[i];
  __syncthreads();
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= total) return;
  // pm = torch.round(pm * 127.5 + 127.5) / 127.5 - 1;  pm = pm / bin_size
  float pm = __fsub_rn(__fdiv_rn(rintf(__fadd_rn(__fmul_rn(p_mean[o], 127.5f), 127.5f)), 127.5f), 1.0f);
  pm = __fdiv_rn(pm, ON_BIN);
  pm_out[o] = pm;
  const float s = fmaxf(expf(__fsub_rn(p_ls[o], ON_LOG_BIN)), 0.11f);          // lower_bound_scale(exp(plogv - log(bin)))
  int k = n_scales - 1;
  for (int t = 0; t < n_scales - 1; ++t) k -= (s <= stab[t]) ? 1 : 0;
  idx[o] = k;
  if (im != nullptr) {
    const float x = __fdiv_rn(__fmul_rn(__fadd_rn(im[o], -0.5f), 2.0f), ON_BIN);
    sym[o] = (int32_t)rintf(__fsub_rn(x, pm));
  }
}

__global__ void __launch_bounds__(256) outnet_decode_kernel(const int32_t* __restrict__ sym, const float* __restrict__ pm,
                                                            float* __restrict__ im_hat, int64_t total) {
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= total) return;
  const float x_hat = __fmul_rn(__fadd_rn((float)sym[o], pm[o]), ON_BIN);
  im_hat[o] = __fadd_rn(__fmul_rn(fminf(fmaxf(x_hat, -1.0f), 1.0f), 0.5f), 0.5f);     // process_output
}

}  // namespace lvae

using namespace lvae;

extern "C" int lvae_nll_output(const float* p_mean, const float* p_logscale, const float* im, float* nll_partial,
                               int B, int chw, void* stream) {
  LVAE_CHECK_ARG(p_mean && p_logscale && im && nll_partial && B > 0 && chw > 0);
  const int np = (chw + ON_T * ON_E - 1) / (ON_T * ON_E);
  nll_output_kernel<<<dim3(np, B), ON_T, 0, (cudaStream_t)stream>>>(p_mean, p_logscale, im, nll_partial, chw, np);
  LVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int lvae_outnet_codec(const float* p_mean, const float* p_logscale, const float* im, const float* scale_table,
                                 int n_scales, float* pm, int32_t* idx, int32_t* sym, int64_t total, void* stream) {
  LVAE_CHECK_ARG(p_mean && p_logscale && scale_table && pm && idx && total > 0 && n_scales >= 1 && n_scales <= 128);
  LVAE_CHECK_ARG((im == nullptr) == (sym == nullptr));
  outnet_codec_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p_mean, p_logscale, im, scale_table, n_scales,
                                                                                         pm, idx, sym, total);
  LVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int lvae_outnet_decode(const int32_t* sym, const float* pm, float* im_hat, int64_t total, void* stream) {
  LVAE_CHECK_ARG(sym && pm && im_hat && total > 0);
  outnet_decode_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(sym, pm, im_hat, total);
  LVAE_CUDA_LAUNCH_CHECK();
  return 0;
}
