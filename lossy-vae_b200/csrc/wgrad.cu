// Weight gradients of the dense layers on the tensor cores (training step):
//   dW[n, k] = sum_p dY[p, n] * X[p, k]        (nn.Linear / 1x1 conv; p runs over the B*H*W pixels)
// is the GEMM of gemm_tc.cu with the pixel index as the contraction dimension.  Both operands are stored pixel-major
// ([P, C] rows), so they are first transposed into K-major 2-plane bf16 operands [C, P] (lvae_split_planes_t, one
// HBM pass each: read 4 B, write 4 B per element); bf16 planes because dY is a gradient (fp32 exponent range, see
// lvae/training.py), two of them = 16 significand bits.  The contraction is split over the persistent grid (split-K,
// ~2 work items per SM) and the partial tiles meet in fp32 atomics.
#include "common.cuh"
#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace lvae {

int gemm_tc_launch_split(const lvae_gemm_desc* d, int split_k, cudaStream_t stream);

// x [P, C] fp32 -> hi / lo bf16 planes [C, P];  P even.  One CTA owns 32 channels x ST_TILES * 64 pixels.
// ACT: planes of gelu(x).  colsum != NULL: colsum[c] += sum over the CTA's pixels of x[p, c] (one atomic per channel).
// tiles: 64-pixel tiles per CTA (<= ST_TILES; fewer when the matrix is small, so that the grid still covers the SMs -- at the
// H/16 ... H/64 stages of a 256 x 256 crop a fixed 8 left 32 CTAs looping 8 times: 15-28 us for a few hundred KB).
constexpr int ST_TILES = 8;
template <bool ACT>
__global__ void __launch_bounds__(256) split_planes_t_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ p0,
                                                             __nv_bfloat16* __restrict__ p1, int64_t P, int C,
                                                             float* __restrict__ colsum, int tiles) {
  __shared__ float tile[64][33];
  __shared__ float red[8][32];
  const int cbase = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  float csum = 0.f;
  for (int it = 0; it < tiles; ++it) {
    const int64_t pbase = ((int64_t)blockIdx.x * tiles + it) * 64;
    if (pbase >= P) break;
    __syncthreads();
    // two rows per step: with ACT the GELU runs in its packed fp32x2 form (same bits as the scalar one, half the issue slots --
    // this kernel is bound by the erf arithmetic, not by memory, when ACT is set)
    for (int r = ty; r < 64; r += 16) {
      const int64_t pa = pbase + r, pb = pbase + r + 8;
      const int c = cbase + tx;
      float2 v = make_float2((pa < P && c < C) ? __ldg(x + pa * C + c) : 0.f, (pb < P && c < C) ? __ldg(x + pb * C + c) : 0.f);
      if (ACT) v = gelu_erf2(v);
      csum += v.x;
      csum += v.y;
      tile[r][tx] = v.x;
      tile[r + 8][tx] = v.y;
    }
    __syncthreads();
    const int64_t pp = pbase + 2 * tx;
    if (pp < P) {
      for (int cc = ty; cc < 32; cc += 8) {
        const int c = cbase + cc;
        if (c >= C) break;
        float2 v = make_float2(tile[2 * tx][cc], tile[2 * tx + 1][cc]);
        const uint32_t hi = split_next<false>(v);
        const uint32_t lo = split_next<false>(v);
        *reinterpret_cast<uint32_t*>(p0 + (int64_t)c * P + pp) = hi;
        *reinterpret_cast<uint32_t*>(p1 + (int64_t)c * P + pp) = lo;
      }
    }
  }
  if (colsum != nullptr) {
    red[ty][tx] = csum;
    __syncthreads();
    if (ty == 0 && cbase + tx < C) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) t += red[i][tx];
      atomicAdd(colsum + cbase + tx, t);
    }
  }
}

// The same transposition for an operand that already exists as two 16-bit planes [P, C] (fp16 or bf16: the A operand of the
// forward GEMM, recomputed at the start of a block's backward): value = hi + lo (exact in fp32), re-split into K-major bf16
// planes [C, P].  Saves the fp32 recomputation of the operand that the fc1 weight gradient would otherwise need.
template <bool F16IN>
__global__ void __launch_bounds__(256) planes_t_kernel(const uint16_t* __restrict__ a0, const uint16_t* __restrict__ a1,
                                                       __nv_bfloat16* __restrict__ p0, __nv_bfloat16* __restrict__ p1,
                                                       int64_t P, int C, int tiles) {
  __shared__ float tile[64][33];
  const int cbase = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  auto cvt = [](uint16_t h) -> float {
    if (F16IN) return __half2float(__ushort_as_half(h));
    return __uint_as_float((uint32_t)h << 16);
  };
  for (int it = 0; it < tiles; ++it) {
    const int64_t pbase = ((int64_t)blockIdx.x * tiles + it) * 64;
    if (pbase >= P) break;
    __syncthreads();
    for (int r = ty; r < 64; r += 8) {
      const int64_t pp = pbase + r;
      const int c = cbase + tx;
      float v = 0.f;
      if (pp < P && c < C) v = __fadd_rn(cvt(__ldg(a0 + pp * C + c)), cvt(__ldg(a1 + pp * C + c)));
      tile[r][tx] = v;
    }
    __syncthreads();
    const int64_t pp = pbase + 2 * tx;
    if (pp < P) {
      for (int cc = ty; cc < 32; cc += 8) {
        const int c = cbase + cc;
        if (c >= C) break;
        float2 v = make_float2(tile[2 * tx][cc], tile[2 * tx + 1][cc]);
        const uint32_t hi = split_next<false>(v);
        const uint32_t lo = split_next<false>(v);
        *reinterpret_cast<uint32_t*>(p0 + (int64_t)c * P + pp) = hi;
        *reinterpret_cast<uint32_t*>(p1 + (int64_t)c * P + pp) = lo;
      }
    }
  }
}

}  // namespace lvae

using namespace lvae;

extern "C" int lvae_planes_transpose(const void* a0, const void* a1, int plane_format, void* p0, void* p1, int64_t P, int C, void* stream) {
  LVAE_CHECK_ARG(a0 && a1 && p0 && p1 && P > 0 && C > 0 && P % 2 == 0);
  LVAE_CHECK_ARG(plane_format == LVAE_PLANES_BF16 || plane_format == LVAE_PLANES_F16);
  int tiles = ST_TILES;
  const int64_t cy = (C + 31) / 32;
  while (tiles > 1 && ((P + 64 * tiles - 1) / (64 * tiles)) * cy < 4 * 148) tiles >>= 1;
  const dim3 grid((unsigned)((P + 64 * tiles - 1) / (64 * tiles)), (unsigned)cy);
  if (plane_format == LVAE_PLANES_F16)
    planes_t_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint16_t*)a0, (const uint16_t*)a1, (__nv_bfloat16*)p0, (__nv_bfloat16*)p1, P, C, tiles);
  else
    planes_t_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint16_t*)a0, (const uint16_t*)a1, (__nv_bfloat16*)p0, (__nv_bfloat16*)p1, P, C, tiles);
  LVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int lvae_split_planes_t_ex(const float* x, void* p0, void* p1, int64_t P, int C, int act, float* colsum, void* stream) {
  LVAE_CHECK_ARG(x && p0 && p1 && P > 0 && C > 0 && P % 2 == 0 && (act == 0 || act == 1));
  int tiles = ST_TILES;
  const int64_t cy = (C + 31) / 32;
  while (tiles > 1 && ((P + 64 * tiles - 1) / (64 * tiles)) * cy < 4 * 148) tiles >>= 1;      // at least ~4 CTAs per SM
  const dim3 grid((unsigned)((P + 64 * tiles - 1) / (64 * tiles)), (unsigned)cy);
  if (act) split_planes_t_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)p0, (__nv_bfloat16*)p1, P, C, colsum, tiles);
  else split_planes_t_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)p0, (__nv_bfloat16*)p1, P, C, colsum, tiles);
  LVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int lvae_split_planes_t(const float* x, void* p0, void* p1, int64_t P, int C, void* stream) {
  return lvae_split_planes_t_ex(x, p0, p1, P, C, 0, nullptr, stream);
}

extern "C" int lvae_gemm_wgrad(const void* dyt_p0, const void* dyt_p1, const void* xt_p0, const void* xt_p1,
                               float* dw, int n_out, int k_in, int64_t P, void* stream) {
  LVAE_CHECK_ARG(dyt_p0 && dyt_p1 && xt_p0 && xt_p1 && dw && n_out > 0 && k_in > 0);
  LVAE_CHECK_ARG(P >= 8 && P % 8 == 0 && P < (1ll << 31));
  cudaStream_t st = (cudaStream_t)stream;
  LVAE_CUDA_CALL(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)n_out * k_in, st));
  lvae_gemm_desc d;
  memset(&d, 0, sizeof(d));
  d.B = 1; d.H = 1; d.W = n_out; d.C0 = (int)P; d.C1 = 0;
  d.ksize = 1; d.stride = 1; d.pad = 0;
  d.N = k_in;
  d.epilogue = LVAE_EPI_BIAS;
  d.out = dw;
  d.precision = LVAE_PREC_BF16X3;
  d.a_planes[0] = dyt_p0; d.a_planes[1] = dyt_p1;
  d.w_planes[0] = xt_p0; d.w_planes[1] = xt_p1;
  return gemm_tc_launch_split(&d, 1, st);
}
