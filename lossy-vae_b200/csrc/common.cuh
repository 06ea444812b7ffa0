// Shared helpers for the lvae_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
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

// torch.nn.GELU() (erf form): 0.5 * x * (1 + erf(x / sqrt(2)))  -- ATen: x * 0.5 * (1 + erf(x * M_SQRT1_2))
__device__ __forceinline__ float gelu_erf(float x) {
  return __fmul_rn(__fmul_rn(x, 0.5f), __fadd_rn(1.0f, erff(__fmul_rn(x, 0.70710678118654752440f))));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace lvae
