// Training-step tail on flat buffers: global-norm gradient clipping + Adam + EMA in two launches.
// Replaces, inside the captured training step, torch.nn.utils.clip_grad_norm_ + torch.optim.Adam + the EMA lerp of the
// reference's loop (lvae/trainer.py:360-377,394-406: scaler.unscale_ -> clip_grad_norm_(max_norm) -> optimizer.step()
// -> ema.update(model)), which as torch foreach ops are ~10 passes over the 907 parameter tensors (374 MB each).
//
//   1. grad_sumsq_kernel   per-block partial sums of g^2 (fixed grid, fixed order: deterministic)
//   2. adam_ema_kernel     every block re-reduces the partials in the same order (-> identical clip coefficient), then
//                          g' = g * min(1, max_norm / (||g|| + 1e-6))                     (clip_grad_norm_)
//                          m = m + (g' - m) (1 - b1);  v = b2 v + (1 - b2) g'^2           (torch.optim.Adam, wd = 0)
//                          p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
//                          e = decay e + (1 - decay) p                                    (ModelEmaV2.update, same roundings)
//   HBM-bound: 20 B read + 16 B written per parameter (36 B; 32 without EMA) against ~56 B for the foreach sequence.
// lr, the step count t and the EMA decay live on the device (the step is replayed as a CUDA graph; the host changes
// them between replays with ordinary copies: learning-rate schedule and EMA warm-up of trainer.py:231-252,373-377).
#include "common.cuh"

namespace lvae {

constexpr int OPT_THREADS = 256;
constexpr int OPT_PARTIALS = 1024;          // blocks of the sum-of-squares pass = partials every Adam block re-reduces

__global__ void __launch_bounds__(OPT_THREADS) grad_sumsq_kernel(const float* __restrict__ g, int64_t n4, int64_t n,
                                                                  double* __restrict__ partial) {
  __shared__ double red[OPT_THREADS / 32];
  double acc = 0.0;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (int64_t i = (int64_t)blockIdx.x * OPT_THREADS + threadIdx.x; i < n4; i += (int64_t)gridDim.x * OPT_THREADS) {
    const float4 v = __ldg(g4 + i);
    acc += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
  }
  if (blockIdx.x == 0) for (int64_t i = n4 * 4 + threadIdx.x; i < n; i += OPT_THREADS) acc += (double)g[i] * g[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < OPT_THREADS / 32; ++w) s += red[w];
    partial[blockIdx.x] = s;
  }
}

struct AdamArgs {
  float* p; const float* g; float* m; float* v; float* ema;
  int64_t n;
  const double* partial; int n_partial;          // sum-of-squares partials (NULL: no clipping)
  float max_norm;
  const float* lr; const float* step; const float* ema_decay;      // device scalars; step = t of THIS update (>= 1)
  double beta1, beta2;                           // as the host's doubles: 1 - beta and beta^t are formed in double, like torch
  float eps;
  float* grad_norm_out;                          // optional: the global gradient norm before clipping
};

__global__ void __launch_bounds__(OPT_THREADS) adam_ema_kernel(const AdamArgs a) {
  __shared__ float s_clip, s_step_size, s_bc2_sqrt;
  if (threadIdx.x == 32) {
    // torch.optim.Adam forms these scalars in Python doubles and rounds once (1 - 0.999f in fp32 is off by 1.3e-5)
    const double t = (double)__ldg(a.step);
    const double bc1 = 1.0 - pow(a.beta1, t), bc2 = 1.0 - pow(a.beta2, t);
    s_step_size = (float)((double)__ldg(a.lr) / bc1);
    s_bc2_sqrt = (float)sqrt(bc2);
  }
  if (a.partial != nullptr) {
    // every block sums the same partials in the same order: one clip coefficient, bit-identical everywhere
    __shared__ double red[OPT_THREADS / 32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < a.n_partial; i += OPT_THREADS) acc += a.partial[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < OPT_THREADS / 32; ++w) s += red[w];
      const float norm = (float)sqrt(s);
      const float coef = a.max_norm / (norm + 1e-6f);            // clip_grad_norm_: clamp(max_norm / (norm + 1e-6), max = 1)
      s_clip = a.max_norm > 0.f ? fminf(coef, 1.0f) : 1.0f;
      if (blockIdx.x == 0 && a.grad_norm_out != nullptr) *a.grad_norm_out = norm;
    }
    __syncthreads();
  } else {
    if (threadIdx.x == 0) s_clip = 1.0f;
    __syncthreads();
  }
  const float clip = s_clip;
  const float step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
  const float w1 = (float)(1.0 - a.beta1), w2 = (float)(1.0 - a.beta2), b2 = (float)a.beta2;
  // ema_decay[0] = decay, [1] = 1 - decay, both rounded from the host's double values as timm's `decay * e + (1. - decay) * m` does
  const float ed = a.ema != nullptr ? __ldg(a.ema_decay) : 0.f, ew = a.ema != nullptr ? __ldg(a.ema_decay + 1) : 0.f;
  const int64_t n4 = a.n >> 2;
  auto upd = [&](float& p, float g, float& m, float& v, float& e) {
    g *= clip;
    m = fmaf(g - m, w1, m);
    v = fmaf(w2 * g, g, b2 * v);
    p -= step_size * (m / (sqrtf(v) / bc2_sqrt + a.eps));
    e = __fadd_rn(__fmul_rn(ed, e), __fmul_rn(ew, p));
  };
  for (int64_t i = (int64_t)blockIdx.x * OPT_THREADS + threadIdx.x; i < n4; i += (int64_t)gridDim.x * OPT_THREADS) {
    float4 p = reinterpret_cast<float4*>(a.p)[i];
    const float4 g = __ldg(reinterpret_cast<const float4*>(a.g) + i);
    float4 m = reinterpret_cast<float4*>(a.m)[i], v = reinterpret_cast<float4*>(a.v)[i];
    float4 e = a.ema ? reinterpret_cast<float4*>(a.ema)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    upd(p.x, g.x, m.x, v.x, e.x); upd(p.y, g.y, m.y, v.y, e.y); upd(p.z, g.z, m.z, v.z, e.z); upd(p.w, g.w, m.w, v.w, e.w);
    reinterpret_cast<float4*>(a.p)[i] = p;
    reinterpret_cast<float4*>(a.m)[i] = m;
    reinterpret_cast<float4*>(a.v)[i] = v;
    if (a.ema) reinterpret_cast<float4*>(a.ema)[i] = e;
  }
  if (blockIdx.x == 0) {
    for (int64_t i = n4 * 4 + threadIdx.x; i < a.n; i += OPT_THREADS) {
      float e = a.ema ? a.ema[i] : 0.f;
      upd(a.p[i], a.g[i], a.m[i], a.v[i], e);
      if (a.ema) a.ema[i] = e;
    }
  }
}

}  // namespace lvae

using namespace lvae;

extern "C" int lvae_optim_scratch_doubles(void) { return OPT_PARTIALS; }

extern "C" int lvae_adam_clip_ema(float* p, const float* g, float* m, float* v, float* ema, int64_t n, double* scratch,
                                  float max_norm, const float* lr, const float* step, const float* ema_decay,
                                  double beta1, double beta2, double eps, float* grad_norm_out, void* stream) {
  LVAE_CHECK_ARG(p && g && m && v && n > 0 && lr && step);
  LVAE_CHECK_ARG(ema == nullptr || ema_decay != nullptr);
  LVAE_CHECK_ARG(((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v | (uintptr_t)ema) % 16 == 0);
  LVAE_CHECK_ARG(max_norm <= 0.f || scratch != nullptr);
  cudaStream_t st = (cudaStream_t)stream;
  const bool clip = max_norm > 0.f || grad_norm_out != nullptr;
  if (clip) {
    LVAE_CHECK_ARG(scratch != nullptr);
    grad_sumsq_kernel<<<OPT_PARTIALS, OPT_THREADS, 0, st>>>(g, n >> 2, n, scratch);
    LVAE_CUDA_LAUNCH_CHECK();
  }
  AdamArgs a;
  a.p = p; a.g = g; a.m = m; a.v = v; a.ema = ema; a.n = n;
  a.partial = clip ? scratch : nullptr; a.n_partial = OPT_PARTIALS; a.max_norm = max_norm;
  a.lr = lr; a.step = step; a.ema_decay = ema_decay; a.beta1 = beta1; a.beta2 = beta2; a.eps = (float)eps;
  a.grad_norm_out = grad_norm_out;
  static int n_sm = 0;
  if (n_sm == 0) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); }
  int64_t blocks = ((n >> 2) + OPT_THREADS - 1) / OPT_THREADS;
  const int64_t cap = (int64_t)n_sm * 8;                       // 8 resident CTAs per SM, grid-stride
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  adam_ema_kernel<<<(unsigned)blocks, OPT_THREADS, 0, st>>>(a);
  LVAE_CUDA_LAUNCH_CHECK();
  return 0;
}
