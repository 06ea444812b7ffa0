// Fused latent-layer kernels: prior transform + quantise + discretised-Gaussian likelihood +
// per-image rate partial sums, one bandwidth-bound pass per latent layer.
//
// Reference arithmetic (fp32, this exact op order -- SURVEY Appendix A):
//   VRLVBlockBase.transform_prior   lvae/models/qarv/model.py:51-53   plogv = softplus(x+2.3)-2.3; pv = exp(plogv)
//   eval branch                     qarv/model.py:95-96 -> CompressAI GaussianConditional.forward:
//       z = rint(qm - pm) + pm ; v = |z - pm| ; s = max(pv, 0.11)
//       P = Phi((.5 - v)/s) - Phi((-.5 - v)/s), Phi(t) = 0.5*(1 + erf(t * 1/sqrt(2))) [td.Normal(0,1).cdf,
//       lvae/models/entropy_coding.py:81-82]; P = max(P, 1e-9); kl = -ln P
//   compress branch                 qarv/model.py:106-108: sym = int(rint(qm-pm)), idx = build_indexes(pv)
//   train branch                    qarv/model.py:91-93 + entropy_coding.py:17-49
//
// Layout: qm / z are [M, zdim] (NHWC), prior is [M, 2*zdim] = (pm | plogv_raw) per position; a thread owns four
// consecutive channels of one position (128-bit loads / stores, 512 B per warp and access), the per-image rate is
// reduced with warp shuffles + one shared-memory step per block.
// Algorithmic bytes per element: read qm, pm, plogv (12 B) + write z (4 B) = 16 B with the rate
// reduced in-kernel (20 B when kl_elem is requested; +8 B for sym/idx).
// The per-image reduction is deterministic: fixed grid, in-block tree, one partial per block.
#include "common.cuh"
#include "latent_math.cuh"

namespace lvae {

constexpr int LT = 256;            // threads per block
constexpr int LE = 4;              // elements per thread

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float t = 0.f;
  if (wid == 0) {
    t = lane < (LT / 32) ? red[lane] : 0.f;
    t = warp_sum(t);
  }
  return t;   // valid in warp 0
}

// VEC = true (zdim % 4 == 0: every registered qarv / rd model): a thread owns 4 consecutive channels of one position --
// 128-bit loads of qm, pm, plogv and 128-bit stores of z / kl_elem, no per-element div / mod, four independent
// transcendental chains in flight per thread; a block still owns LT * LE consecutive elements of one image, so the
// per-block partial sums land where they did.  VEC = false: the scalar layout for odd channel counts (qres z_dims 14, 10).
template <int MODE, bool VEC>
__global__ void __launch_bounds__(LT) latent_kernel(
    const float* __restrict__ qm, const float* __restrict__ prior, const float* __restrict__ noise,
    const float* __restrict__ table, int n_scales,
    float* __restrict__ z, float* __restrict__ kl_partial, float* __restrict__ kl_elem,
    int32_t* __restrict__ sym, int32_t* __restrict__ idx, int hw, int zdim, int kl_stride, int cdf_kind) {
  __shared__ float red[LT / 32];
  __shared__ float stab[64];
  if (MODE == 0 && idx != nullptr) {
    for (int i = threadIdx.x; i < n_scales && i < 64; i += LT) stab[i] = table[i];
    __syncthreads();
  }
  const int b = blockIdx.y;
  const int per_img = hw * zdim;
  float local = 0.f;
  if (VEC) {
    const int i = blockIdx.x * (LT * LE) + threadIdx.x * 4;        // first of 4 consecutive elements (same position)
    if (i < per_img) {
      const int pos = i / zdim, c = i - pos * zdim;
      const int64_t m = (int64_t)b * hw + pos;
      const float4 q4 = __ldg(reinterpret_cast<const float4*>(qm + m * zdim + c));
      const float4 pm4 = __ldg(reinterpret_cast<const float4*>(prior + m * 2 * zdim + c));
      const float4 pl4 = __ldg(reinterpret_cast<const float4*>(prior + m * 2 * zdim + zdim + c));
      float4 n4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (MODE == 1) n4 = __ldg(reinterpret_cast<const float4*>(noise + m * zdim + c));
      const float q[4] = {q4.x, q4.y, q4.z, q4.w}, pm[4] = {pm4.x, pm4.y, pm4.z, pm4.w};
      const float pl[4] = {pl4.x, pl4.y, pl4.z, pl4.w}, nz[4] = {n4.x, n4.y, n4.z, n4.w};
      float zz[4], kl[4], r[4], sc[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) latent_elem<MODE>(q[e], pm[e], pl[e], nz[e], cdf_kind, zz[e], kl[e], r[e], sc[e]);
      *reinterpret_cast<float4*>(z + m * zdim + c) = make_float4(zz[0], zz[1], zz[2], zz[3]);
      if (kl_elem != nullptr) *reinterpret_cast<float4*>(kl_elem + m * zdim + c) = make_float4(kl[0], kl[1], kl[2], kl[3]);
      if (MODE == 0 && sym != nullptr) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int64_t o = ((int64_t)b * zdim + c + e) * hw + pos;     // NCHW order for the coder
          sym[o] = (int32_t)r[e];
          idx[o] = scale_index(sc[e], stab, n_scales);
        }
      }
      local = ((kl[0] + kl[1]) + kl[2]) + kl[3];
    }
  } else {
#pragma unroll
    for (int e = 0; e < LE; ++e) {
      const int i = (blockIdx.x * LE + e) * LT + threadIdx.x;     // element within the image, (pos, c)
      if (i < per_img) {
        const int pos = i / zdim, c = i - pos * zdim;
        const int64_t m = (int64_t)b * hw + pos;
        float zz, kl, r, sc;
        latent_elem<MODE>(qm[m * zdim + c], prior[m * 2 * zdim + c], prior[m * 2 * zdim + zdim + c],
                          MODE == 1 ? noise[m * zdim + c] : 0.f, cdf_kind, zz, kl, r, sc);
        if (MODE == 0 && sym != nullptr) {
          const int64_t o = ((int64_t)b * zdim + c) * hw + pos;     // NCHW order for the coder
          sym[o] = (int32_t)r;
          idx[o] = scale_index(sc, stab, n_scales);
        }
        z[m * zdim + c] = zz;
        if (kl_elem != nullptr) kl_elem[m * zdim + c] = kl;
        local += kl;
      }
    }
  }
  const float t = block_sum(local, red);
  if (threadIdx.x == 0) kl_partial[(int64_t)b * kl_stride + blockIdx.x] = t;
}

// Backward of the training branch: kl = -gaussian_log_prob_mass(pm, sigma, z), z = qm + noise,
// sigma = exp(softplus(raw + 2.3) - 2.3).  torch.where routes the gradient through the selected branch only
// (lvae/models/entropy_coding.py:17-26; the unselected branch's gradient is finite, so it contributes exactly 0):
//   mass branch (mass > 1e-6), u, l = (z +- 1/2 - pm) / sigma, phi = standard normal pdf:
//     d lnP/dz = (phi(u) - phi(l)) / (sigma P),  d lnP/dpm = -d lnP/dz,  d lnP/dsigma = -(u phi(u) - l phi(l)) / (sigma P)
//   tail branch: lnP = Normal(pm, sigma).log_prob(z):  d/dz = -(z - pm)/sigma^2,  d/dsigma = (z - pm)^2/sigma^3 - 1/sigma
// and d sigma/d raw = sigma * sigmoid(raw + 2.3) (softplus with torch's threshold 20).
__global__ void __launch_bounds__(LT) latent_train_bwd_kernel(
    const float* __restrict__ qm, const float* __restrict__ prior, const float* __restrict__ noise,
    const float* __restrict__ dz, const float* __restrict__ dkl_elem, float dkl_scale,
    float* __restrict__ dqm, float* __restrict__ dprior, int64_t total, int zdim) {
  const int64_t i = (int64_t)blockIdx.x * LT + threadIdx.x;
  if (i >= total) return;
  const int64_t m = i / zdim; const int c = (int)(i - m * zdim);
  const float q = qm[i];
  const float pm = prior[m * 2 * zdim + c];
  const float raw = prior[m * 2 * zdim + zdim + c];
  const float pv = prior_scale(raw);
  const float zz = __fadd_rn(q, noise[i]);
  const float rcp = __frcp_rn(pv);
  const float u = __fmul_rn(__fsub_rn(__fadd_rn(zz, 0.5f), pm), rcp);
  const float l = __fmul_rn(__fsub_rn(__fsub_rn(zz, 0.5f), pm), rcp);
  const float cu = __fmul_rn(0.5f, __fadd_rn(1.0f, erf_torch_cpu(__fdiv_rn(u, 1.4142135623730951f))));
  const float cl = __fmul_rn(0.5f, __fadd_rn(1.0f, erf_torch_cpu(__fdiv_rn(l, 1.4142135623730951f))));
  const float mass = __fsub_rn(cu, cl);
  float dlp_dz, dlp_ds;
  if (mass > 1e-6f) {
    const float pu = 0.3989422804014327f * expf(-0.5f * u * u), pl = 0.3989422804014327f * expf(-0.5f * l * l);
    const float inv = rcp / mass;
    dlp_dz = (pu - pl) * inv;
    dlp_ds = -(u * pu - l * pl) * inv;
  } else {
    const float d = __fsub_rn(zz, pm);
    dlp_dz = -d * rcp * rcp;
    dlp_ds = d * d * rcp * rcp * rcp - rcp;
  }
  const float g = dkl_elem ? dkl_elem[i] : dkl_scale;       // dL/dkl, kl = -lnP
  const float x = __fadd_rn(raw, 2.3f);
  const float sig = x > 20.0f ? 1.0f : 1.0f / (1.0f + expf(-x));
  const float dzt = (dz ? dz[i] : 0.f) - g * dlp_dz;
  dqm[i] = dzt;
  dprior[m * 2 * zdim + c] = g * dlp_dz;                    // dL/dpm = -g * dlnP/dpm = +g * dlnP/dz
  dprior[m * 2 * zdim + zdim + c] = -g * dlp_ds * pv * sig;
}

__global__ void __launch_bounds__(LT) prior_index_kernel(
    const float* __restrict__ prior, const float* __restrict__ table, int n_scales,
    int32_t* __restrict__ idx, int hw, int zdim) {
  __shared__ float stab[64];
  for (int i = threadIdx.x; i < n_scales && i < 64; i += LT) stab[i] = table[i];
  __syncthreads();
  const int b = blockIdx.y;
  const int i = blockIdx.x * LT + threadIdx.x;
  if (i >= hw * zdim) return;
  const int pos = i / zdim, c = i - pos * zdim;
  const int64_t m = (int64_t)b * hw + pos;
  const float s = fmaxf(prior_scale(prior[m * 2 * zdim + zdim + c]), 0.11f);
  int k = n_scales - 1;
  for (int t = 0; t < n_scales - 1; ++t) k -= (s <= stab[t]) ? 1 : 0;
  idx[((int64_t)b * zdim + c) * hw + pos] = k;
}

__global__ void __launch_bounds__(LT) dequant_kernel(
    const int32_t* __restrict__ sym, const float* __restrict__ prior, float* __restrict__ z, int hw, int zdim) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * LT + threadIdx.x;
  if (i >= hw * zdim) return;
  const int pos = i / zdim, c = i - pos * zdim;
  const int64_t m = (int64_t)b * hw + pos;
  const float pm = prior[m * 2 * zdim + c];
  z[m * zdim + c] = __fadd_rn((float)sym[((int64_t)b * zdim + c) * hw + pos], pm);
}

__global__ void __launch_bounds__(LT) sample_kernel(
    const float* __restrict__ prior, const float* __restrict__ randn, const float* __restrict__ unif, float t,
    float* __restrict__ z, int64_t M, int zdim) {
  const int64_t i = (int64_t)blockIdx.x * LT + threadIdx.x;
  if (i >= M * zdim) return;
  const int64_t m = i / zdim; const int c = (int)(i - m * zdim);
  const float pm = prior[m * 2 * zdim + c];
  const float pv = prior_scale(prior[m * 2 * zdim + zdim + c]);
  // z = pm + pv * randn * t + unif * t       (qarv/model.py:100)
  z[i] = __fadd_rn(__fadd_rn(pm, __fmul_rn(__fmul_rn(pv, randn[i]), t)), __fmul_rn(unif[i], t));
}

// ---- rd model (continuous Gaussian posterior; lvae/models/rd/model.py:27-49,162-227) ----------------------------
// linear_sqrt (rd/model.py:27-39): sign(x) |x|^(1 - tanh(|x|)/2) for |x| <= 6, sign(x) sqrt(|x| + 1e-8) beyond
__device__ __forceinline__ float rd_linear_sqrt(float x) {
  const float a = fabsf(x);
  if (a == 0.0f) return x;
  float v;
  if (a <= 6.0f) v = powf(a, __fsub_rn(1.0f, __fmul_rn(0.5f, tanhf(a))));
  else v = sqrtf(__fadd_rn(a, 1e-8f));
  return copysignf(v, x);
}
// F.softplus(v, beta = ln 2, threshold = 12) (rd/model.py:162-165)
__device__ __forceinline__ float rd_std_smooth(float v) {
  const float beta = 0.6931471805599453f;
  const float vb = __fmul_rn(v, beta);
  return vb > 12.0f ? v : __fdiv_rn(log1pf(expf(vb)), beta);
}

// post / prior: [M, 2*zdim] = (mean_raw | std_raw) per position; noise [M, zdim] ~ N(0,1) supplied by the caller
__global__ void __launch_bounds__(LT) rd_latent_kernel(
    const float* __restrict__ post, const float* __restrict__ prior, const float* __restrict__ noise,
    float* __restrict__ z, float* __restrict__ kl_partial, float* __restrict__ kl_elem, int hw, int zdim, int kl_stride) {
  __shared__ float red[LT / 32];
  const int b = blockIdx.y;
  const int per_img = hw * zdim;
  float local = 0.f;
#pragma unroll
  for (int e = 0; e < LE; ++e) {
    const int i = (blockIdx.x * LE + e) * LT + threadIdx.x;
    if (i < per_img) {
      const int pos = i / zdim, c = i - pos * zdim;
      const int64_t m = (int64_t)b * hw + pos;
      const float qm = rd_linear_sqrt(post[m * 2 * zdim + c]);
      const float qv = rd_std_smooth(post[m * 2 * zdim + zdim + c]);
      const float pm = rd_linear_sqrt(prior[m * 2 * zdim + c]);
      const float pv = rd_std_smooth(prior[m * 2 * zdim + zdim + c]);
      // gaussian_kl(qm, qv, pm, pv) = -0.5 + log(pv) - log(qv) + 0.5 * (qv^2 + (qm - pm)^2) / pv^2, left to right
      const float d = __fsub_rn(qm, pm);
      const float t2 = __fsub_rn(__fadd_rn(-0.5f, logf(pv)), logf(qv));
      const float t5 = __fdiv_rn(__fmul_rn(0.5f, __fadd_rn(__fmul_rn(qv, qv), __fmul_rn(d, d))), __fmul_rn(pv, pv));
      const float kl = __fadd_rn(t2, t5);
      z[m * zdim + c] = __fadd_rn(qm, __fmul_rn(qv, noise[m * zdim + c]));      // rd/model.py:213
      if (kl_elem != nullptr) kl_elem[m * zdim + c] = kl;
      local += kl;
    }
  }
  const float t = block_sum(local, red);
  if (threadIdx.x == 0) kl_partial[(int64_t)b * kl_stride + blockIdx.x] = t;
}

__global__ void __launch_bounds__(LT) rd_sample_kernel(
    const float* __restrict__ prior, const float* __restrict__ randn, float t, float* __restrict__ z, int64_t M, int zdim) {
  const int64_t i = (int64_t)blockIdx.x * LT + threadIdx.x;
  if (i >= M * zdim) return;
  const int64_t m = i / zdim; const int c = (int)(i - m * zdim);
  const float pm = rd_linear_sqrt(prior[m * 2 * zdim + c]);
  const float pv = rd_std_smooth(prior[m * 2 * zdim + zdim + c]);
  z[i] = __fadd_rn(pm, __fmul_rn(__fmul_rn(pv, randn[i]), t));                    // rd/model.py:216
}

}  // namespace lvae

using namespace lvae;

extern "C" int lvae_latent_num_partials(int hw, int zdim) {
  const int per = LT * LE;
  return (hw * zdim + per - 1) / per;
}

extern "C" int lvae_latent_eval(const float* qm, const float* prior, const float* scale_table, int n_scales,
                                float* z, float* kl_partial, int kl_stride, float* kl_elem,
                                int32_t* sym, int32_t* idx, int B, int hw, int zdim, int cdf_kind, void* stream) {
  LVAE_CHECK_ARG(qm && prior && z && kl_partial && B > 0 && hw > 0 && zdim > 0);
  LVAE_CHECK_ARG(cdf_kind == LVAE_CDF_NORMAL || cdf_kind == LVAE_CDF_ERFC);
  LVAE_CHECK_ARG((sym == nullptr) == (idx == nullptr));
  LVAE_CHECK_ARG(sym == nullptr || (scale_table != nullptr && n_scales >= 1 && n_scales <= 64));
  const int np = lvae_latent_num_partials(hw, zdim);
  LVAE_CHECK_ARG(kl_stride >= np);
  const bool vec = zdim % 4 == 0 && ((uintptr_t)qm | (uintptr_t)prior | (uintptr_t)z | (uintptr_t)kl_elem) % 16 == 0;
  if (vec) latent_kernel<0, true><<<dim3(np, B), LT, 0, (cudaStream_t)stream>>>(qm, prior, nullptr, scale_table, n_scales,
                                                                                 z, kl_partial, kl_elem, sym, idx, hw, zdim, kl_stride, cdf_kind);
  else latent_kernel<0, false><<<dim3(np, B), LT, 0, (cudaStream_t)stream>>>(qm, prior, nullptr, scale_table, n_scales,
                                                                              z, kl_partial, kl_elem, sym, idx, hw, zdim, kl_stride, cdf_kind);
  LVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int lvae_latent_train(const float* qm, const float* prior, const float* noise,
                                 float* z, float* kl_partial, int kl_stride, float* kl_elem,
                                 int B, int hw, int zdim, void* stream) {
  LVAE_CHECK_ARG(qm && prior && noise && z && kl_partial && B > 0 && hw > 0 && zdim > 0);
  const int np = lvae_latent_num_partials(hw, zdim);
  LVAE_CHECK_ARG(kl_stride >= np);
  const bool vec = zdim % 4 == 0 && ((uintptr_t)qm | (uintptr_t)prior | (uintptr_t)noise | (uintptr_t)z | (uintptr_t)kl_elem) % 16 == 0;
  if (vec) latent_kernel<1, true><<<dim3(np, B), LT, 0, (cudaStream_t)stream>>>(qm, prior, noise, nullptr, 0,
                                                                                 z, kl_partial, kl_elem, nullptr, nullptr, hw, zdim, kl_stride, 0);
  else latent_kernel<1, false><<<dim3(np, B), LT, 0, (cudaStream_t)stream>>>(qm, prior, noise, nullptr, 0,
                                                                              z, kl_partial, kl_elem, nullptr, nullptr, hw, zdim, kl_stride, 0);
  LVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int lvae_latent_train_bwd(const float* qm, const float* prior, const float* noise,
                                     const float* dz, const float* dkl_elem, float dkl_scale,
                                     float* dqm, float* dprior, int B, int hw, int zdim, void* stream) {
  LVAE_CHECK_ARG(qm && prior && noise && dqm && dprior && B > 0 && hw > 0 && zdim > 0);
  const int64_t total = (int64_t)B * hw * zdim;
  latent_train_bwd_kernel<<<(unsigned)((total + LT - 1) / LT), LT, 0, (cudaStream_t)stream>>>(
      qm, prior, noise, dz, dkl_elem, dkl_scale, dqm, dprior, total, zdim);
  LVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int lvae_latent_prior_index(const float* prior, const float* scale_table, int n_scales,
                                       int32_t* idx, int B, int hw, int zdim, void* stream) {
  LVAE_CHECK_ARG(prior && scale_table && idx && n_scales >= 1 && n_scales <= 64 && B > 0 && hw > 0 && zdim > 0);
  prior_index_kernel<<<dim3((hw * zdim + LT - 1) / LT, B), LT, 0, (cudaStream_t)stream>>>(prior, scale_table, n_scales, idx, hw, zdim);
  LVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int lvae_latent_dequant(const int32_t* sym, const float* prior, float* z,
                                   int B, int hw, int zdim, void* stream) {
  LVAE_CHECK_ARG(sym && prior && z && B > 0 && hw > 0 && zdim > 0);
  dequant_kernel<<<dim3((hw * zdim + LT - 1) / LT, B), LT, 0, (cudaStream_t)stream>>>(sym, prior, z, hw, zdim);
  LVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int lvae_latent_sample(const float* prior, const float* randn, const float* unif, float t,
                                  float* z, int B, int hw, int zdim, void* stream) {
  LVAE_CHECK_ARG(prior && randn && unif && z && B > 0 && hw > 0 && zdim > 0);
  const int64_t M = (int64_t)B * hw;
  sample_kernel<<<(unsigned)((M * zdim + LT - 1) / LT), LT, 0, (cudaStream_t)stream>>>(prior, randn, unif, t, z, M, zdim);
  LVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int lvae_rd_latent(const float* post, const float* prior, const float* noise,
                              float* z, float* kl_partial, int kl_stride, float* kl_elem,
                              int B, int hw, int zdim, void* stream) {
  LVAE_CHECK_ARG(post && prior && noise && z && kl_partial && B > 0 && hw > 0 && zdim > 0);
  const int np = lvae_latent_num_partials(hw, zdim);
  LVAE_CHECK_ARG(kl_stride >= np);
  rd_latent_kernel<<<dim3(np, B), LT, 0, (cudaStream_t)stream>>>(post, prior, noise, z, kl_partial, kl_elem, hw, zdim, kl_stride);
  LVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int lvae_rd_sample(const float* prior, const float* randn, float t, float* z,
                              int B, int hw, int zdim, void* stream) {
  LVAE_CHECK_ARG(prior && randn && z && B > 0 && hw > 0 && zdim > 0);
  const int64_t M = (int64_t)B * hw;
  rd_sample_kernel<<<(unsigned)((M * zdim + LT - 1) / LT), LT, 0, (cudaStream_t)stream>>>(prior, randn, t, z, M, zdim);
  LVAE_CUDA_LAUNCH_CHECK();
  return 0;
}
