// Arithmetic of one latent element (prior transform, quantise, discretised-Gaussian likelihood, table index), shared by the
// stand-alone fused latent kernels (latent.cu) and the latent epilogue of the posterior convolution (gemm_tc.cu): ONE definition,
// so both produce the same bits.  Reference op order: see the header of latent.cu.
#pragma once
#include "common.cuh"

namespace lvae {

__device__ __forceinline__ float softplus_torch(float x) {   // beta=1, threshold=20
  return x > 20.0f ? x : log1pf(expf(x));
}
__device__ __forceinline__ float prior_scale(float plogv_raw) {
  const float plogv = __fsub_rn(softplus_torch(__fadd_rn(plogv_raw, 2.3f)), 2.3f);
  return expf(plogv);
}
// td.Normal(0,1).cdf(t) = 0.5 * (1 + erf((t - 0) * (1/1) / sqrt(2)))
__device__ __forceinline__ float std_normal_cdf(float t) {
  const float a = __fdiv_rn(t, 1.4142135623730951f);
  return __fmul_rn(0.5f, __fadd_rn(1.0f, erf_torch_cpu(a)));
}

// CompressAI GaussianConditional._standardized_cumulative: 0.5 * erfc(-(2^-0.5) * t)
__device__ __forceinline__ float std_normal_cdf_erfc(float t) {
  return __fmul_rn(0.5f, erfcf(__fmul_rn(-0.70710678118654752440f, t)));
}

// One latent element: z, kl (and for the coder the integer symbol r and the clamped scale s).  mode 0: eval, 1: train.
template <int MODE>
__device__ __forceinline__ void latent_elem(float q, float pm, float plogv_raw, float nz, int cdf_kind,
                                            float& zz, float& kl, float& r, float& s) {
  const float pv = prior_scale(plogv_raw);
  if (MODE == 0) {
    r = rintf(__fsub_rn(q, pm));                           // torch.round: half to even
    zz = __fadd_rn(r, pm);
    const float v = fabsf(__fsub_rn(zz, pm));
    s = fmaxf(pv, 0.11f);
    const float tu = __fdiv_rn(__fsub_rn(0.5f, v), s), tl = __fdiv_rn(__fsub_rn(-0.5f, v), s);
    const float up = cdf_kind ? std_normal_cdf_erfc(tu) : std_normal_cdf(tu);
    const float lo = cdf_kind ? std_normal_cdf_erfc(tl) : std_normal_cdf(tl);
    const float P = fmaxf(__fsub_rn(up, lo), 1e-9f);
    kl = -logf(P);
  } else {
    r = 0.f; s = pv;
    zz = __fadd_rn(q, nz);
    // td.Normal(pm, pv).cdf(x) = 0.5*(1+erf((x-pm)*(1/pv)/sqrt(2)))
    const float rcp = __frcp_rn(pv);
    const float cu = __fmul_rn(0.5f, __fadd_rn(1.0f, erf_torch_cpu(__fdiv_rn(__fmul_rn(__fsub_rn(__fadd_rn(zz, 0.5f), pm), rcp), 1.4142135623730951f))));
    const float cl = __fmul_rn(0.5f, __fadd_rn(1.0f, erf_torch_cpu(__fdiv_rn(__fmul_rn(__fsub_rn(__fsub_rn(zz, 0.5f), pm), rcp), 1.4142135623730951f))));
    const float mass = __fsub_rn(cu, cl);
    float lp;
    if (mass > 1e-6f) {
      lp = logf(fmaxf(mass, 1e-8f));
    } else {
      // Normal.log_prob: -((x-mu)^2)/(2 var) - log(scale) - log(sqrt(2 pi));  + log(bin_size=1) = 0
      const float d = __fsub_rn(zz, pm);
      const float var = __fmul_rn(pv, pv);
      lp = __fsub_rn(__fsub_rn(__fdiv_rn(-__fmul_rn(d, d), __fmul_rn(2.0f, var)), logf(pv)), 0.9189385332046727f);
      lp = __fadd_rn(lp, 0.0f);
    }
    kl = -lp;
  }
}

__device__ __forceinline__ int scale_index(float s, const float* stab, int n_scales) {
  int k = n_scales - 1;
  for (int t = 0; t < n_scales - 1; ++t) k -= (s <= stab[t]) ? 1 : 0;
  return k;
}

}  // namespace lvae
