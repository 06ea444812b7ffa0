// Small operators around the hot kernels: lambda embedding, tiny-M linear layers (all AdaLN
// projections in one launch), image pre/post-processing and deterministic reductions.
#include "common.cuh"
#include <cuda_bf16.h>
#include <stdarg.h>
#include <math.h>
#include <stdlib.h>

namespace lvae {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}

bool pdl_enabled() {
  static int on = -1;
  // opt-in: measured on B200 (profiles/r2_bench_notes.md) the whole-step graph replays no faster with programmatic edges
  // (559 vs 566 images/s) -- a persistent GEMM CTA owns its SM's shared memory, so a dependent CTA cannot become resident
  // before the last tile anyway -- so plain stream order stays the default
  if (on < 0) { const char* e = getenv("LVAE_PDL"); on = (e && atoi(e) == 1) ? 1 : 0; }
  return on != 0;
}

// common.py:101-107 + qarv/model.py:275-279
__global__ void sinusoid_kernel(const float* __restrict__ lmb, const float* __restrict__ freqs,
                                float* __restrict__ emb0, int B, int dim, float period, float log_max) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = dim / 2;
  if (i >= B * half) return;
  const int b = i / half, f = i - b * half;
  // lmb_input = torch.log(lmb) * period / math.log(MAX_LMB)   (left to right, fp32)
  const float scaled = __fdiv_rn(__fmul_rn(logf(lmb[b]), period), log_max);
  const float arg = __fmul_rn(scaled, freqs[f]);
  emb0[(int64_t)b * dim + f] = cosf(arg);
  emb0[(int64_t)b * dim + half + f] = sinf(arg);
}

// Small-batch linear layer: out[b, n] = act_out(sum_k act_in(x[b, k]) w[n, k] + bias[n]).  The (activated) input
// [B, K] is staged once per block in shared memory -- with act_in = GELU (all 90 AdaLN projections in one launch,
// N ~ 63 k) the old one-warp-per-column kernel evaluated the GELU N * K * B = 129 M times and was compute-bound at
// 0.4 TB/s of weight traffic.  Each warp then streams SL_COLS weight rows; per-lane K order and the warp reduction
// are unchanged, so results are bit-identical to the previous kernel.
constexpr int SL_COLS = 8;                  // output columns per warp
template <int BCHUNK>
__global__ void __launch_bounds__(256) small_linear_kernel(
    const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
    float* __restrict__ out, int B, int K, int N, int act_in, int act_out) {
  extern __shared__ float sl_x[];            // [B][K]
  for (int i = threadIdx.x; i < B * K; i += blockDim.x) {
    const float v = __ldg(x + i);
    sl_x[i] = act_in ? gelu_erf(v) : v;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int n0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * SL_COLS;
  for (int n = n0; n < n0 + SL_COLS && n < N; ++n) {
    const float* wr = w + (int64_t)n * K;
    for (int b0 = 0; b0 < B; b0 += BCHUNK) {
      float acc[BCHUNK];
#pragma unroll
      for (int i = 0; i < BCHUNK; ++i) acc[i] = 0.f;
      for (int k = lane * 4; k < K; k += 128) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(wr + k));
#pragma unroll
        for (int i = 0; i < BCHUNK; ++i) {
          if (b0 + i < B) {
            const float4 xv = *reinterpret_cast<const float4*>(sl_x + (b0 + i) * K + k);
            acc[i] = fmaf(xv.x, wv.x, acc[i]); acc[i] = fmaf(xv.y, wv.y, acc[i]);
            acc[i] = fmaf(xv.z, wv.z, acc[i]); acc[i] = fmaf(xv.w, wv.w, acc[i]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < BCHUNK; ++i) {
        const float s = warp_sum(acc[i]);
        if (lane == 0 && b0 + i < B) {
          float v = bias ? __fadd_rn(s, bias[n]) : s;
          if (act_out) v = gelu_erf(v);
          out[(int64_t)(b0 + i) * N + n] = v;
        }
      }
    }
  }
}

// qarv/model.py:213-222 preprocess_input fused with the space-to-depth gather of patch_downsample(r)
__global__ void image_to_patches_kernel(const float* __restrict__ im, float* __restrict__ a,
                                        int B, int H, int W, int r, float shift, float scale) {
  const int Ho = H / r, Wo = W / r, K = 3 * r * r;
  const int64_t total = (int64_t)B * Ho * Wo * K;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int k = (int)(i % K); const int64_t m = i / K;
  const int c = k % 3; const int q = k / 3; const int ii = q / r, jj = q - ii * r;
  const int wo = (int)(m % Wo); const int64_t t = m / Wo; const int ho = (int)(t % Ho); const int b = (int)(t / Ho);
  const float v = im[(((int64_t)b * 3 + c) * H + ho * r + ii) * W + wo * r + jj];
  a[i] = __fmul_rn(__fadd_rn(v, shift), scale);
}

constexpr int DT = 256, DE = 8;
__global__ void __launch_bounds__(DT) image_distortion_kernel(
    const float* __restrict__ x_hat, const float* __restrict__ im, float* __restrict__ im_hat,
    float* __restrict__ p_tgt, float* __restrict__ p_im, int chw, int nparts) {
  __shared__ float red[2][DT / 32];
  const int b = blockIdx.y;
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int e = 0; e < DE; ++e) {
    const int i = (blockIdx.x * DE + e) * DT + threadIdx.x;
    if (i < chw) {
      const int64_t o = (int64_t)b * chw + i;
      const float xh = x_hat[o], v = im[o];
      const float tgt = __fmul_rn(__fadd_rn(v, -0.5f), 2.0f);          // preprocess_target
      const float d1 = __fsub_rn(xh, tgt);
      s1 = fmaf(d1, d1, s1);
      const float ih = __fadd_rn(__fmul_rn(fminf(fmaxf(xh, -1.0f), 1.0f), 0.5f), 0.5f);  // process_output
      if (im_hat) im_hat[o] = ih;
      const float d2 = __fsub_rn(ih, v);
      s2 = fmaf(d2, d2, s2);
    }
  }
  s1 = warp_sum(s1); s2 = warp_sum(s2);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { red[0][wid] = s1; red[1][wid] = s2; }
  __syncthreads();
  if (wid == 0) {
    float a = lane < DT / 32 ? red[0][lane] : 0.f, c = lane < DT / 32 ? red[1][lane] : 0.f;
    a = warp_sum(a); c = warp_sum(c);
    if (lane == 0) { p_tgt[(int64_t)b * nparts + blockIdx.x] = a; p_im[(int64_t)b * nparts + blockIdx.x] = c; }
  }
}

__global__ void broadcast_bias_kernel(const float* __restrict__ bias, float* __restrict__ out, int64_t total4, int C4) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  reinterpret_cast<float4*>(out)[i] = __ldg(reinterpret_cast<const float4*>(bias) + (i % C4));
}

__global__ void sum_partials_kernel(const float* __restrict__ partial, float* __restrict__ out, int n, int cols) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int j = 0; j < cols; ++j) s += (double)partial[(int64_t)i * cols + j];
  out[i] = (float)s;
}

// Loss assembly of VariableRateLossyVAE.forward (qarv/model.py:338-358) from the deterministic
// partial sums written by the latent and distortion kernels.  One block; warp w handles images
// w, w+nwarps, ...; lanes stride over the partial columns; double accumulation in a fixed order.
__global__ void __launch_bounds__(256) rd_finalize_kernel(
    const float* __restrict__ kl_partial, int kl_stride, int kl_cols,
    const float* __restrict__ p_tgt, const float* __restrict__ p_im, int np,
    const float* __restrict__ lmb, int B, float ndims, float* __restrict__ stats) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float* kl_img = stats + 4; float* mse_img = stats + 4 + B; float* sqim_img = stats + 4 + 2 * B;
  for (int b = wid; b < B; b += nw) {
    double a = 0.0, t = 0.0, u = 0.0;
    for (int j = lane; j < kl_cols; j += 32) a += (double)kl_partial[(int64_t)b * kl_stride + j];
    for (int j = lane; j < np; j += 32) { t += (double)p_tgt[(int64_t)b * np + j]; u += (double)p_im[(int64_t)b * np + j]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o); t += __shfl_xor_sync(0xffffffffu, t, o); u += __shfl_xor_sync(0xffffffffu, u, o);
    }
    if (lane == 0) {
      kl_img[b] = __fdiv_rn((float)a, ndims);          // nats per dimension
      mse_img[b] = __fdiv_rn((float)t, ndims);         // mse vs (im-.5)*2, mean over C,H,W
      sqim_img[b] = (float)u;                          // sum of squared error of clamp(x_hat)*.5+.5 vs im
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double loss = 0.0, kl = 0.0, mse = 0.0, sq = 0.0;
    for (int b = 0; b < B; ++b) {
      const float lb = __fadd_rn(kl_img[b], __fmul_rn(lmb[b], mse_img[b]));
      loss += (double)lb; kl += (double)kl_img[b]; mse += (double)mse_img[b]; sq += (double)sqim_img[b];
    }
    stats[0] = (float)(loss / B); stats[1] = (float)(kl / B); stats[2] = (float)(mse / B);
    stats[3] = (float)(sq / ((double)B * (double)ndims));
  }
}

__global__ void split_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ p0,
                                  __nv_bfloat16* __restrict__ p1, __nv_bfloat16* __restrict__ p2, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = x[i];
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  p0[i] = h;
  if (p1) {
    const float r1 = __fsub_rn(v, __bfloat162float(h));              // exact
    const __nv_bfloat16 m = __float2bfloat16_rn(r1);
    p1[i] = m;
    if (p2) p2[i] = __float2bfloat16_rn(__fsub_rn(r1, __bfloat162float(m)));
  }
}

__global__ void pad_channels_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t total, int C, int Cp) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int64_t m = i / Cp; const int c = (int)(i - m * Cp);
  dst[i] = c < C ? src[m * C + c] : 0.f;
}

// two elements per thread: planes of (x * scale) in either 16-bit format (split_next, common.cuh)
__global__ void split_planes_kernel(const float* __restrict__ x, uint32_t* __restrict__ p0, uint32_t* __restrict__ p1,
                                    uint32_t* __restrict__ p2, int64_t n2, int f16, float scale, int act) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n2) return;
  float2 v = __ldg(reinterpret_cast<const float2*>(x) + i);
  if (act) { v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); }
  v.x = __fmul_rn(v.x, scale); v.y = __fmul_rn(v.y, scale);
  p0[i] = split_next(v, f16 != 0);
  if (p1) {
    p1[i] = split_next(v, f16 != 0);
    if (p2) p2[i] = split_next(v, f16 != 0);
  }
}

}  // namespace lvae

using namespace lvae;

extern "C" int lvae_version(void) { return 101; }
extern "C" const char* lvae_last_error(void) { return lvae::g_err; }

extern "C" int lvae_lmb_sinusoid(const float* lmb, const float* freqs, float* emb0, int B, int dim,
                                 float period, float max_lmb, void* stream) {
  // freqs: [dim/2] host-generated table max_period^(-linspace(0,1,dim/2)) (fp32, made by torch on the host)
  LVAE_CHECK_ARG(lmb && freqs && emb0 && B > 0 && dim > 0 && dim % 2 == 0 && max_lmb > 1.f);
  const int total = B * (dim / 2);
  sinusoid_kernel<<<(total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(lmb, freqs, emb0, B, dim, period, (float)log((double)max_lmb));
  LVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int lvae_small_linear(const float* x, const float* w, const float* bias, float* out,
                                 int B, int K, int N, int act_in, int act_out, void* stream) {
  LVAE_CHECK_ARG(x && w && out && B > 0 && K > 0 && K % 4 == 0 && N > 0);
  const int warps = 8, cols = warps * SL_COLS;
  const size_t smem = (size_t)B * K * sizeof(float);
  LVAE_CHECK_ARG(smem <= 160 * 1024);
  if (smem > 48 * 1024) LVAE_CUDA_CALL(cudaFuncSetAttribute(small_linear_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  small_linear_kernel<8><<<(N + cols - 1) / cols, warps * 32, smem, (cudaStream_t)stream>>>(x, w, bias, out, B, K, N, act_in, act_out);
  LVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int lvae_image_to_patches(const float* im, float* a, int B, int H, int W, int r,
                                     float shift, float scale, void* stream) {
  LVAE_CHECK_ARG(im && a && B > 0 && r > 0 && H % r == 0 && W % r == 0);
  const int64_t total = (int64_t)B * H * W * 3;
  image_to_patches_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(im, a, B, H, W, r, shift, scale);
  LVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

// GPU-side input pipeline of the training step: RandomCrop + RandomHorizontalFlip + ToTensor of
// lvae/datasets/image.py:45-56 on a batch of decoded uint8 images that is already on the device.  src [B, 3, Hs, Ws]
// uint8 (NCHW), per-image crop origin (y0, x0) and flip flag -> out [B, 3, crop, crop] fp32 in [0, 1] (x / 255, the value
// torchvision's to_tensor produces).  Pure data movement: 1 byte read + 4 bytes written per element.
namespace lvae {
__global__ void __launch_bounds__(256) crop_flip_kernel(const uint8_t* __restrict__ src, const int32_t* __restrict__ y0,
                                                        const int32_t* __restrict__ x0, const uint8_t* __restrict__ flip,
                                                        float* __restrict__ out, int B, int Hs, int Ws, int crop) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)B * 3 * crop * crop;
  if (i >= total) return;
  const int x = (int)(i % crop); int64_t t = i / crop;
  const int y = (int)(t % crop); t /= crop;
  const int c = (int)(t % 3); const int b = (int)(t / 3);
  const int sx = x0[b] + (flip[b] ? crop - 1 - x : x), sy = y0[b] + y;
  out[i] = __fdiv_rn((float)src[(((int64_t)b * 3 + c) * Hs + sy) * Ws + sx], 255.0f);
}
}  // namespace lvae

extern "C" int lvae_crop_flip_u8(const void* src, const int32_t* y0, const int32_t* x0, const void* flip, float* out,
                                 int B, int Hs, int Ws, int crop, void* stream) {
  LVAE_CHECK_ARG(src && y0 && x0 && flip && out && B > 0 && crop > 0 && crop <= Hs && crop <= Ws);
  const int64_t total = (int64_t)B * 3 * crop * crop;
  lvae::crop_flip_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const uint8_t*)src, y0, x0, (const uint8_t*)flip, out, B, Hs, Ws, crop);
  LVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int lvae_image_num_partials(int chw) { return (chw + DT * DE - 1) / (DT * DE); }

extern "C" int lvae_image_distortion(const float* x_hat, const float* im, float* im_hat,
                                     float* sq_target_partial, float* sq_im_partial, int B, int chw, void* stream) {
  LVAE_CHECK_ARG(x_hat && im && sq_target_partial && sq_im_partial && B > 0 && chw > 0);
  const int np = lvae_image_num_partials(chw);
  image_distortion_kernel<<<dim3(np, B), DT, 0, (cudaStream_t)stream>>>(x_hat, im, im_hat, sq_target_partial, sq_im_partial, chw, np);
  LVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int lvae_broadcast_bias(const float* bias, float* out, int64_t M, int C, void* stream) {
  LVAE_CHECK_ARG(bias && out && M > 0 && C > 0 && C % 4 == 0);
  const int64_t total4 = M * (C / 4);
  broadcast_bias_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(bias, out, total4, C / 4);
  LVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int lvae_pad_channels(const float* src, float* dst, int64_t M, int C, int Cp, void* stream) {
  LVAE_CHECK_ARG(src && dst && M > 0 && C > 0 && Cp >= C);
  const int64_t total = M * Cp;
  pad_channels_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, dst, total, C, Cp);
  LVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int lvae_sum_partials(const float* partial, float* out, int n, int cols, void* stream) {
  LVAE_CHECK_ARG(partial && out && n > 0 && cols > 0);
  sum_partials_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(partial, out, n, cols);
  LVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int lvae_split_bf16(const float* x, void* p0, void* p1, void* p2, int64_t n, void* stream) {
  LVAE_CHECK_ARG(x && p0 && n > 0 && (p2 == nullptr || p1 != nullptr));
  split_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      x, (__nv_bfloat16*)p0, (__nv_bfloat16*)p1, (__nv_bfloat16*)p2, n);
  LVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int lvae_split_planes(const float* x, void* p0, void* p1, void* p2, int64_t n, int plane_format, float scale,
                                 void* stream) {
  LVAE_CHECK_ARG(x && p0 && n > 0 && n % 2 == 0 && (p2 == nullptr || p1 != nullptr));
  LVAE_CHECK_ARG(plane_format == LVAE_PLANES_BF16 || plane_format == LVAE_PLANES_F16);
  LVAE_CHECK_ARG(scale > 0.f);
  split_planes_kernel<<<(unsigned)((n / 2 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      x, (uint32_t*)p0, (uint32_t*)p1, (uint32_t*)p2, n / 2, plane_format == LVAE_PLANES_F16, scale, 0);
  LVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int lvae_gelu_split_planes(const float* x, void* p0, void* p1, void* p2, int64_t n, int plane_format, void* stream) {
  LVAE_CHECK_ARG(x && p0 && n > 0 && n % 2 == 0 && (p2 == nullptr || p1 != nullptr));
  LVAE_CHECK_ARG(plane_format == LVAE_PLANES_BF16 || plane_format == LVAE_PLANES_F16);
  split_planes_kernel<<<(unsigned)((n / 2 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      x, (uint32_t*)p0, (uint32_t*)p1, (uint32_t*)p2, n / 2, plane_format == LVAE_PLANES_F16, 1.0f, 1);
  LVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int lvae_rd_finalize(const float* kl_partial, int kl_stride, int kl_cols,
                                const float* sq_target_partial, const float* sq_im_partial, int np,
                                const float* lmb, int B, int64_t ndims, float* stats, void* stream) {
  LVAE_CHECK_ARG(kl_partial && sq_target_partial && sq_im_partial && lmb && stats);
  LVAE_CHECK_ARG(B > 0 && kl_cols > 0 && kl_stride >= kl_cols && np > 0 && ndims > 0);
  rd_finalize_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(kl_partial, kl_stride, kl_cols, sq_target_partial,
                                                          sq_im_partial, np, lmb, B, (float)ndims, stats);
  LVAE_CUDA_LAUNCH_CHECK();
  return 0;
}
