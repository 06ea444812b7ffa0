// Shared helpers for the lvae_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <utility>
#include "lvae_b200.h"

namespace lvae {

void set_error(const char* fmt, ...);

#define LVAE_CHECK_ARG(cond)                                                        \
  do {                                                                              \
    if (!(cond)) {                                                                  \
      ::lvae::set_error("%s:%d: bad argument: %s", __FILE__, __LINE__, #cond);      \
      return LVAE_E_BADARG;                                                         \
    }                                                                               \
  } while (0)

#define LVAE_CUDA_LAUNCH_CHECK()                                                    \
  do {                                                                              \
    cudaError_t e__ = cudaGetLastError();                                           \
    if (e__ != cudaSuccess) {                                                       \
      ::lvae::set_error("%s:%d: CUDA: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return (int)e__;                                                              \
    }                                                                               \
  } while (0)

#define LVAE_CUDA_CALL(x)                                                           \
  do {                                                                              \
    cudaError_t e__ = (x);                                                          \
    if (e__ != cudaSuccess) {                                                       \
      ::lvae::set_error("%s:%d: CUDA: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return (int)e__;                                                              \
    }                                                                               \
  } while (0)

// ---- programmatic dependent launch (PDL).  A kernel launched through launch_pdl() may start while its predecessor in
// the stream (or in the captured graph) is still draining: its CTAs become resident as SMs free up and run their
// prologue (barrier init, TMEM allocation, tensor-map prefetch) early, then block in pdl_wait() until the predecessor
// has completed and its memory is visible.  RULE: every kernel launched with launch_pdl() calls pdl_wait() before its
// first global-memory access that could alias the predecessor's reads or writes, and pdl_trigger() as early as possible.
// Opt-in with LVAE_PDL=1 (environment); without the launch attribute the device calls are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// erf as ATen's CPU GELU kernel evaluates it (the parity oracle is the reference on the CPU): at::vec::Vectorized
// <float>::erf(), i.e. Abramowitz-Stegun 7.1.26 in fp32 with FMAs (torch/include/ATen/cpu/vec/vec512/
// vec512_float.h:269-299; |error| <= 1.5e-7 -- larger than an fp32 ulp, so matching the formula, not the true
// erf, is what keeps hidden activations within rounding noise of the oracle).
// Carries -t instead of t (every sign flip is exact) and takes e^{-x^2} from an inlined range reduction + ex2.approx
// (<= 2 ulp, like expf) so that the packed form below can mirror it instruction by instruction.
//
// torch.nn.GELU() (erf form) on CPU: x * 0.5 * (1 + erf(x * M_SQRT1_2))  (ATen native/cpu/Activation.cpp GeluKernelImpl).
// gelu_erf() and gelu_erf2() return the SAME BITS for every input (tests/test_gpu_kernels.py::
// test_gelu_scalar_and_packed_forms_agree_bitwise): which of the two an element goes through depends on where it falls in a
// tile (full 32-column chunks are packed, edge chunks scalar), and the tile shape may depend on the batch size.
__device__ __forceinline__ float gelu_erf(float x) {
  const float xk = __fmul_rn(x, 0.70710678118654752440f);
  const float a = fabsf(xk);
  // t = 1 / (1 + p|x|): the argument is a normal number >= 1, so the correctly rounded reciprocal is the MUFU
  // approximation plus one Newton step (== _mm512_div_ps(1, .) on the host)
  const float d = fmaf(0.3275911f, a, 1.0f);
  float tn;                                                                    // -1/d
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(tn) : "f"(-d));
  tn = fmaf(tn, fmaf(d, tn, 1.0f), tn);
  float r = fmaf(1.061405429f, tn, 1.453152027f);                              // -(p5 t + p4)
  r = fmaf(r, tn, 1.421413741f);
  r = fmaf(r, tn, 0.284496736f);
  r = fmaf(r, tn, 0.254829592f);
  float y = fmaxf(__fmul_rn(xk, -xk), -87.0f);
  const float z = fmaf(y, 1.4426950408889634f, 12583039.0f);                   // low mantissa bits: n + 127
  const float nf = __fadd_rn(z, -12583039.0f);
  float u = fmaf(y, 1.4426950216293335f, -nf);
  u = fmaf(y, 1.925963033500011e-08f, u);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(u));
  e = __fmul_rn(e, __uint_as_float(__float_as_uint(z) << 23));
  const float res = fmaf(__fmul_rn(e, tn), r, 1.0f);                           // 1 - e t r
  return __fmul_rn(__fmul_rn(x, 0.5f), __fadd_rn(1.0f, copysignf(res, xk)));
}

// ---- packed fp32x2 arithmetic (sm_100): each lane is an IEEE op, i.e. bit-identical to the scalar instruction, at
// half the issue slots.  Used where the CUDA cores, not memory, bound a kernel (GELU epilogue, depthwise conv).
__device__ __forceinline__ uint64_t f2_as_u64(float2 v) { return *reinterpret_cast<uint64_t*>(&v); }
__device__ __forceinline__ float2 u64_as_f2(uint64_t v) { return *reinterpret_cast<float2*>(&v); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2_as_u64(a)), "l"(f2_as_u64(b)), "l"(f2_as_u64(c)));
  return u64_as_f2(d);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  uint64_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_as_u64(a)), "l"(f2_as_u64(b)));
  return u64_as_f2(d);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_as_u64(a)), "l"(f2_as_u64(b)));
  return u64_as_f2(d);
}
__device__ __forceinline__ float2 splat2(float v) { return make_float2(v, v); }

// gelu_erf on two values: the same instructions in their packed fp32x2 forms around the two MUFU ops -- the same bits.
__device__ __forceinline__ float2 gelu_erf2(float2 x) {
  const float2 xk = mul2(x, splat2(0.70710678118654752440f));
  const float2 a = make_float2(fabsf(xk.x), fabsf(xk.y));
  const float2 d = fma2(splat2(0.3275911f), a, splat2(1.0f));                 // 1 + p|x|  (>= 1)
  float2 tn;                                                                   // -1/d
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(tn.x) : "f"(-d.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(tn.y) : "f"(-d.y));
  tn = fma2(tn, fma2(d, tn, splat2(1.0f)), tn);                                // Newton step: correctly rounded -1/d
  float2 r = fma2(splat2(1.061405429f), tn, splat2(1.453152027f));             // -(p5 t + p4)
  r = fma2(r, tn, splat2(1.421413741f));                                       //  (..) t + p3
  r = fma2(r, tn, splat2(0.284496736f));                                       // -((..) t + p2)
  r = fma2(r, tn, splat2(0.254829592f));                                       //  (..) t + p1
  // e = exp(-xk^2)
  float2 y = mul2(xk, make_float2(-xk.x, -xk.y));
  y.x = fmaxf(y.x, -87.0f); y.y = fmaxf(y.y, -87.0f);
  const float2 z = fma2(y, splat2(1.4426950408889634f), splat2(12583039.0f)); // low mantissa bits: n + 127
  const float2 nf = add2(z, splat2(-12583039.0f));
  float2 u = fma2(y, splat2(1.4426950216293335f), make_float2(-nf.x, -nf.y));
  u = fma2(y, splat2(1.925963033500011e-08f), u);
  float2 e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(u.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(u.y));
  e = mul2(e, make_float2(__uint_as_float(__float_as_uint(z.x) << 23), __uint_as_float(__float_as_uint(z.y) << 23)));
  const float2 res = fma2(mul2(e, tn), r, splat2(1.0f));                       // 1 - e t r
  const float2 erf = make_float2(copysignf(res.x, xk.x), copysignf(res.y, xk.y));
  return mul2(mul2(x, splat2(0.5f)), add2(splat2(1.0f), erf));
}

// torch.erf on CPU float tensors (what td.Normal.cdf calls): a <= 0.55-ulp erf that saturates to +-1 from
// |x| >= 3.8325069 (measured against this image's torch 2.11 / MKL build over every float in [1e-3, 4.5]:
// 97 % of results equal the correctly rounded value, the rest are off by one ulp only where the true value is
// within 0.05 ulp of a rounding midpoint).  CUDA's erff (2 ulp) saturates only from 3.919, which moves tail
// likelihoods by whole quanta of 2^-25; 1 - erfcf(|x|) is correctly rounded except near midpoints.
__device__ __forceinline__ float erf_torch_cpu(float x) {
  const float a = fabsf(x);
  float r;
  if (a >= 3.8325069f) r = 1.0f;
  else if (a > 1.5f) r = __fsub_rn(1.0f, erfcf(a));   // erfcf: 4 ulp relative -> < 0.3 ulp of erf here
  else r = erff(a);
  return copysignf(r, x);
}

// ---- 16-bit operand planes of the tensor-core GEMM (lvae_plane_format).  pack2<F16>(a, b) rounds two fp32 values to
// bf16 (F16 = false) or fp16 (F16 = true; saturating, so an out-of-range activation becomes +-65504 instead of inf)
// and returns them as one 32-bit word, a in the low half; unpack2 is the exact inverse widening.
template <bool F16> __device__ __forceinline__ uint32_t pack2(float a, float b) {
  uint32_t d;
  if (F16) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
  else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
  return d;
}
template <bool F16> __device__ __forceinline__ float2 unpack2(uint32_t d) {
  if (F16) {
    float2 r;
    asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tcvt.f32.f16 %0, lo;\n\tcvt.f32.f16 %1, hi;\n\t}"
        : "=f"(r.x), "=f"(r.y) : "r"(d));
    return r;
  }
  return make_float2(__uint_as_float(d << 16), __uint_as_float(d & 0xffff0000u));
}
// plane i of the pair (v.x, v.y); v is replaced by the exact residual v - plane for the next plane
template <bool F16> __device__ __forceinline__ uint32_t split_next(float2& v) {
  const uint32_t w = pack2<F16>(v.x, v.y);
  const float2 f = unpack2<F16>(w);
  v.x = __fsub_rn(v.x, f.x);
  v.y = __fsub_rn(v.y, f.y);
  return w;
}
__device__ __forceinline__ uint32_t split_next(float2& v, bool f16) { return f16 ? split_next<true>(v) : split_next<false>(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace lvae
