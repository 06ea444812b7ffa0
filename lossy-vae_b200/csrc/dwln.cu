// Depthwise kxk conv + LayerNorm(C, eps 1e-6) + AdaLN (or affine LN), NHWC fp32 in, fp32 and/or 16-bit planes out.
// Reference: ConvNeXtBlockAdaLN.forward, lvae/models/common.py:145-152 (conv_dw -> permute -> norm ->
// x*(1+scale)+shift); qresvae MyConvNeXtBlock (affine LayerNorm, no AdaLN) qresvae/model.py:163-182.
//
// Balanced between HBM (algorithmic bytes per position: read 4C, write 4C fp32 or 2C per 16-bit plane) and the fp32
// FMA pipe (49 MACs per output at k = 7).  One CTA of 4 warps owns an 8 x 8 tile of output pixels of one image and
// 64 channels (128 for the rd model's 640 / 768); a thread-block cluster of C / 64 CTAs covers all channels:
//   * the (8+k-1)^2 x 64 halo is one TMA box load (cp.async.bulk.tensor.4d over the [C, W, H, B] view;
//     coordinates outside the image are zero-filled by the hardware = the conv's zero padding): no fill loop, no
//     index arithmetic; at 49 KB per CTA four CTAs share an SM, so loads, convolution, the cluster exchange and the
//     stores of different tiles overlap (measured: 4 warps x 4 CTAs beats 8 x 2 and 6 x 3 by 10-15 %);
//   * warp w convolves output rows 2w and 2w+1 together, lane l owns channels {2l, 2l+1}: rolling over the k+1 input
//     rows, every shared-memory row and every filter row is loaded once and feeds both outputs -- 161 loads per 784
//     packed FFMA2 instead of 294, which moves the bound from the L1/shared pipe (62 % busy before) to the FMA pipe;
//   * LayerNorm: per-CTA two-pass statistics in registers (16-value halving warp reduction: 16 shuffles instead of
//     80), then Chan's parallel-variance combination of the CL partial (sum, M2) pairs, always in rank order, so every
//     CTA of the cluster -- and every launch -- agrees bit for bit.  The partials are PUSHED: every CTA writes its pair
//     into each peer's shared memory with st.async (complete_tx on the peer's mbarrier) and waits on its own mbarrier
//     for CL x 512 bytes.  No cluster barrier on the critical path: r1's barrier.cluster release / acquire pair and the
//     exit barrier were MEMBAR.ALL.GPU + ERRBAR waits, 25 % of the kernel's stall samples (ncu source view);
//   * modulation (AdaLN or affine) and the split into 16-bit planes for the tensor-core GEMM that consumes them.
#include "common.cuh"
#include <cuda_bf16.h>
#include <cuda.h>
#include <cooperative_groups.h>
#include <mutex>
#include <stdlib.h>
#include <type_traits>

namespace lvae {

#ifndef LVAE_DW_WARPS
#define LVAE_DW_WARPS 4
#endif
constexpr int DW_WARPS = LVAE_DW_WARPS; // warps per CTA
constexpr int DW_TW = 8;                // output tile: 8 columns x 2 * DW_WARPS rows, warp w owns rows 2w and 2w + 1
constexpr int DW_TH = 2 * DW_WARPS;
constexpr int DW_CH = 64;               // channels per chunk (lane l owns channels 2l, 2l + 1 of the chunk)
constexpr int DW_PIX = 2 * DW_TW;       // pixels per warp

__device__ __forceinline__ uint32_t dw_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dw_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void dw_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void dw_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void dw_tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// remote (distributed shared memory) 8-byte store that credits 8 bytes to an mbarrier of the destination CTA
__device__ __forceinline__ void dw_st_async_v2(uint32_t remote_addr, float a, float b, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];"
               ::"r"(remote_addr), "f"(a), "f"(b), "r"(remote_bar) : "memory");
}
__device__ __forceinline__ uint32_t dw_mapa(uint32_t addr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
// packed fp32x2 FMA (sm_100): both lanes are IEEE fma, i.e. bit-identical to two fmaf() at half the issue slots
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d)
      : "l"(*reinterpret_cast<const uint64_t*>(&a)), "l"(*reinterpret_cast<const uint64_t*>(&b)),
        "l"(*reinterpret_cast<const uint64_t*>(&c)));
  return *reinterpret_cast<float2*>(&d);
}

// Sum of 16 per-lane values over the 32 lanes of a warp in 16 shuffles (instead of 16 x 5): at every butterfly step a
// lane keeps the half of the values its partner does not.  Lane l ends with the total of value (l >> 1) & 15; the
// additions are the same pairs in the same order as the xor-butterfly warp_sum(), so the totals are bit-identical.
__device__ __forceinline__ float warp_sum16(const float (&v)[DW_PIX], int lane) {
  float a[8], b[4], c[2];
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float keep = h16 ? v[i + 8] : v[i], send = h16 ? v[i] : v[i + 8];
    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float keep = h8 ? a[i + 4] : a[i], send = h8 ? a[i] : a[i + 4];
    b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float keep = h4 ? b[i + 2] : b[i], send = h4 ? b[i] : b[i + 2];
    c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  const float keep = h2 ? c[1] : c[0], send = h2 ? c[0] : c[1];
  float d = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  d += __shfl_xor_sync(0xffffffffu, d, 1);
  return d;
}

// One CTA (8 warps) = one 16 x 8 pixel tile x (NJ * 64) channels; a cluster of CL CTAs covers all C = NJ * 64 * CL
// channels of the tile and exchanges LayerNorm partial statistics through distributed shared memory.
template <int NJ, int KS, int CL>
__global__ void __launch_bounds__(32 * DW_WARPS) dwln_kernel(
    const __grid_constant__ CUtensorMap x_map, const float* __restrict__ dw_w, const float* __restrict__ dw_b,
    const float* __restrict__ ada, int64_t ada_stride, int64_t ada_off,
    const float* __restrict__ ln_w, const float* __restrict__ ln_b,
    float* __restrict__ y, __nv_bfloat16* __restrict__ y0, __nv_bfloat16* __restrict__ y1, __nv_bfloat16* __restrict__ y2,
    int f16, int H, int W, int tiles_x, int tiles_y, int nbatch, int l2_ahead) {
  constexpr int C = NJ * DW_CH * CL, NLOC = NJ * DW_CH, PAD = (KS - 1) / 2;
  constexpr int HW_ = DW_TW + KS - 1, HH_ = DW_TH + KS - 1;                  // halo tile width / height
  constexpr int CHUNK_FLOATS = HH_ * HW_ * DW_CH;
  constexpr int NBUF = NJ > 1 ? 2 : 1;
  extern __shared__ __align__(128) float dw_smem[];                          // [NBUF][HH_][HW_][64]
  __shared__ __align__(8) uint64_t dw_bar[3];                                       // [0], [1]: halo chunks; [2]: LayerNorm partials of the cluster
  __shared__ __align__(16) float2 ln_recv[CL][DW_WARPS * DW_PIX];                   // (sum, M2) per tile pixel, written by CTA r of the cluster
  __shared__ __align__(16) float ln_loc[DW_WARPS][2][DW_PIX];                       // per warp: broadcast scratch
  const int tid = threadIdx.x, lane = tid & 31, wrow = tid >> 5;
  const int crank = (CL > 1) ? (int)(blockIdx.x % CL) : 0;                   // cluster dims (CL,1,1): rank == blockIdx.x % CL
  const int cbase = crank * NLOC;                                            // first channel of this CTA
  int t = blockIdx.x / CL;
  const int tx = t % tiles_x; t /= tiles_x;
  const int ty = t % tiles_y; const int b = t / tiles_y;
  const int h0 = ty * DW_TH, w0 = tx * DW_TW;

  if (tid == 0) {
    dw_mbar_init(dw_smem_u32(&dw_bar[0]), 1);
    dw_mbar_init(dw_smem_u32(&dw_bar[1]), 1);
    if (CL > 1) {
      // armed for the whole exchange right away: every CTA of the cluster (this one included) sends 8 bytes per pixel
      dw_mbar_init(dw_smem_u32(&dw_bar[2]), 1);
      dw_mbar_expect_tx(dw_smem_u32(&dw_bar[2]), (uint32_t)(CL * DW_WARPS * DW_PIX * 8));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // peers may signal dw_bar[2] only once it exists: arrive now, wait before the first remote store.  fence.mbarrier_init
  // above is what publishes the initialisation, so the arrive can be RELAXED (the release form costs a MEMBAR.ALL.GPU +
  // ERRBAR: 9 % of the kernel's stall samples when it stood here)
  if (CL > 1) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  auto load_chunk = [&](int j, int buf) {                                    // one thread
    const uint32_t bar = dw_smem_u32(&dw_bar[buf]);
    dw_mbar_expect_tx(bar, (uint32_t)(CHUNK_FLOATS * 4));
    dw_tma_load_4d(dw_smem_u32(dw_smem + buf * CHUNK_FLOATS), &x_map, bar, cbase + j * DW_CH, w0 - PAD, h0 - PAD, b);
  };

  // PDL: barrier set-up and the cluster arrive above overlap the predecessor's tail; the halo load below reads its output
  pdl_trigger();
  pdl_wait();
  // res[j][p]: conv output of pixel p = r * 8 + s (r = 0, 1: rows 2 wrow + r; s: column) for this lane's 2 channels
  float2 res[NJ][DW_PIX];
  if (tid == 0) {
    load_chunk(0, 0);
    // pull the halo of the tile a later CTA of this SM will want into L2 (one wave of resident CTAs ahead)
    if (l2_ahead > 0) {
      int t2 = (int)(blockIdx.x / CL) + l2_ahead;
      const int tx2 = t2 % tiles_x; t2 /= tiles_x;
      const int ty2 = t2 % tiles_y; const int b2 = t2 / tiles_y;
      if (b2 < nbatch)
        for (int j = 0; j < NJ; ++j)
          asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global [%0, {%1, %2, %3, %4}];"
                       ::"l"(&x_map), "r"(cbase + j * DW_CH), "r"(tx2 * DW_TW - PAD), "r"(ty2 * DW_TH - PAD), "r"(b2) : "memory");
    }
  }
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    // buffer (j+1)&1 was last read in iteration j-1, which ended with __syncthreads()
    if (NJ > 1 && tid == 0 && j + 1 < NJ) load_chunk(j + 1, (j + 1) & 1);
    const int c = cbase + j * DW_CH + lane * 2;
    // bias and the first filter row are in flight while the halo arrives; row i + 1 is fetched while row i is used
    const float2 bias = __ldg(reinterpret_cast<const float2*>(dw_b + c));
    float2 wnext[KS];
#pragma unroll
    for (int kx = 0; kx < KS; ++kx) wnext[kx] = __ldg(reinterpret_cast<const float2*>(dw_w + kx * C + c));
    dw_mbar_wait(dw_smem_u32(&dw_bar[j & (NBUF - 1)]), (uint32_t)((j >> 1) & 1));
    // the peers' matching arrive was the first thing they did (cluster CTAs start together): this returns at once, and
    // from here on every peer's dw_bar[2] is known to exist
    if (CL > 1 && j == 0) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    const float* tile = dw_smem + (j & (NBUF - 1)) * CHUNK_FLOATS;
#pragma unroll
    for (int p = 0; p < DW_PIX; ++p) res[j][p] = bias;
    // rolling over the KS + 1 halo rows that feed output rows o1 = 2 wrow (tap row ky = i) and o2 = o1 + 1 (ky = i - 1):
    // every input row is read from shared memory once, every filter row is loaded once and used for both outputs;
    // per output the taps still accumulate in (ky, kx) order, i.e. the sums are those of the one-row-per-warp kernel
    float2 wprev[KS];
#pragma unroll
    for (int i = 0; i <= KS; ++i) {
      const float* row = tile + ((2 * wrow + i) * HW_) * DW_CH + lane * 2;
      float2 xv[HW_];
#pragma unroll
      for (int q = 0; q < HW_; ++q) xv[q] = *reinterpret_cast<const float2*>(row + q * DW_CH);
      float2 wcur[KS];
#pragma unroll
      for (int kx = 0; kx < KS; ++kx) wcur[kx] = wnext[kx];
      if (i + 1 < KS) {
#pragma unroll
        for (int kx = 0; kx < KS; ++kx) wnext[kx] = __ldg(reinterpret_cast<const float2*>(dw_w + ((i + 1) * KS + kx) * C + c));
      }
#pragma unroll
      for (int kx = 0; kx < KS; ++kx) {
        if (i >= 1) {
#pragma unroll
          for (int s = 0; s < DW_TW; ++s) res[j][DW_TW + s] = ffma2(xv[s + kx], wprev[kx], res[j][DW_TW + s]);
        }
      }
#pragma unroll
      for (int kx = 0; kx < KS; ++kx) {
        if (i < KS) {
#pragma unroll
          for (int s = 0; s < DW_TW; ++s) res[j][s] = ffma2(xv[s + kx], wcur[kx], res[j][s]);
        }
      }
#pragma unroll
      for (int kx = 0; kx < KS; ++kx) wprev[kx] = wcur[kx];
    }
    if (NJ > 1) __syncthreads();           // everyone is done with this buffer before chunk j+2 overwrites it
  }

  // modulation coefficients v * mul + add (mul = (1 + scale) | gamma, add = shift | beta): the loads are issued here so
  // that their latency hides behind the statistics and the cluster exchange below
  float2 mod_mul[NJ], mod_add[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int c = cbase + j * DW_CH + lane * 2;
    if (ln_w != nullptr) {
      mod_mul[j] = __ldg(reinterpret_cast<const float2*>(ln_w + c));
      mod_add[j] = __ldg(reinterpret_cast<const float2*>(ln_b + c));
    } else {
      const float* e = ada + (int64_t)b * ada_stride + ada_off + c;
      mod_add[j] = __ldg(reinterpret_cast<const float2*>(e));
      mod_mul[j] = __ldg(reinterpret_cast<const float2*>(e + C));        // scale; 1 + scale is formed at the use
    }
  }

  // ---- LayerNorm statistics: two-pass over this CTA's channels (registers), Chan's combination across the cluster
  namespace cg = cooperative_groups;
  float v16[DW_PIX];
#pragma unroll
  for (int p = 0; p < DW_PIX; ++p) {
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) sum += res[j][p].x + res[j][p].y;
    v16[p] = sum;
  }
  float tot = warp_sum16(v16, lane);                           // lane l: sum of pixel (l >> 1) & 15 over NLOC channels
  if ((lane & 1) == 0) ln_loc[wrow][0][lane >> 1] = tot;
  __syncwarp();
  float mean[DW_PIX], rstd[DW_PIX];
#pragma unroll
  for (int p4 = 0; p4 < DW_PIX; p4 += 4) {
    const float4 s4 = *reinterpret_cast<const float4*>(&ln_loc[wrow][0][p4]);
    mean[p4] = s4.x; mean[p4 + 1] = s4.y; mean[p4 + 2] = s4.z; mean[p4 + 3] = s4.w;     // local sums for now
  }
#pragma unroll
  for (int p = 0; p < DW_PIX; ++p) {
    const float mloc = mean[p] * (1.0f / NLOC);
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const float dx = res[j][p].x - mloc, dy = res[j][p].y - mloc;
      sq = fmaf(dx, dx, sq); sq = fmaf(dy, dy, sq);
    }
    v16[p] = sq;
  }
  const float m2 = warp_sum16(v16, lane);                      // centred second moment about the LOCAL mean
  if (CL == 1) {
    __syncwarp();
    if ((lane & 1) == 0) ln_loc[wrow][1][lane >> 1] = m2;
    __syncwarp();
#pragma unroll
    for (int p4 = 0; p4 < DW_PIX; p4 += 4) {
      const float4 q4 = *reinterpret_cast<const float4*>(&ln_loc[wrow][1][p4]);
      rstd[p4] = q4.x; rstd[p4 + 1] = q4.y; rstd[p4 + 2] = q4.z; rstd[p4 + 3] = q4.w;
    }
#pragma unroll
    for (int p = 0; p < DW_PIX; ++p) {
      mean[p] = mean[p] * (1.0f / C);
      rstd[p] = 1.0f / sqrtf(rstd[p] * (1.0f / C) + 1e-6f);
    }
  } else {
    if ((lane & 1) == 0) {
      // push this CTA's (sum, M2) of pixel (lane >> 1) to every CTA of the cluster, itself included
      const uint32_t slot = dw_smem_u32(&ln_recv[crank][wrow * DW_PIX + (lane >> 1)]);
      const uint32_t bar = dw_smem_u32(&dw_bar[2]);
#pragma unroll
      for (int r = 0; r < CL; ++r) dw_st_async_v2(dw_mapa(slot, (uint32_t)r), tot, m2, dw_mapa(bar, (uint32_t)r));
    }
    dw_mbar_wait(dw_smem_u32(&dw_bar[2]), 0);
    if (lane < DW_PIX) {
      // pixel `lane` of this warp: combine the CL partial (sum, M2) pairs in rank order (identical in every CTA)
      float s_r[CL], q_r[CL];
#pragma unroll
      for (int r = 0; r < CL; ++r) {
        const float2 pr = ln_recv[r][wrow * DW_PIX + lane];
        s_r[r] = pr.x; q_r[r] = pr.y;
      }
      float S = 0.f;
#pragma unroll
      for (int r = 0; r < CL; ++r) S += s_r[r];
      const float mu = S * (1.0f / C);
      float M2 = 0.f;
#pragma unroll
      for (int r = 0; r < CL; ++r) {
        const float dm = s_r[r] * (1.0f / NLOC) - mu;
        M2 += fmaf((float)NLOC * dm, dm, q_r[r]);
      }
      ln_loc[wrow][0][lane] = mu;
      ln_loc[wrow][1][lane] = 1.0f / sqrtf(M2 * (1.0f / C) + 1e-6f);
    }
    __syncwarp();
#pragma unroll
    for (int p4 = 0; p4 < DW_PIX; p4 += 4) {
      const float4 s4 = *reinterpret_cast<const float4*>(&ln_loc[wrow][0][p4]);
      const float4 q4 = *reinterpret_cast<const float4*>(&ln_loc[wrow][1][p4]);
      mean[p4] = s4.x; mean[p4 + 1] = s4.y; mean[p4 + 2] = s4.z; mean[p4 + 3] = s4.w;
      rstd[p4] = q4.x; rstd[p4 + 1] = q4.y; rstd[p4 + 2] = q4.z; rstd[p4 + 3] = q4.w;
    }
  }

  // ---- modulation + output: one specialised store loop per (output kind, interior / border tile) so that the
  //      per-pixel work carries no pointer tests or bounds checks (they were 18 % of the kernel's instructions)
  const bool interior = (h0 + DW_TH <= H) && (w0 + DW_TW <= W);
  const int64_t o00 = (((int64_t)b * H + h0 + 2 * wrow) * W + w0) * C;        // pixel (row 2 wrow, column 0) of the tile
  auto emit = [&](auto mode_c, auto interior_c) {
    constexpr int MODE = decltype(mode_c)::value;          // 0: fp32 | 1-3: bf16 planes | 4: two fp16 planes | 5: one fp16 plane
    constexpr bool INTERIOR = decltype(interior_c)::value;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int c = cbase + j * DW_CH + lane * 2;
      const float2 add = mod_add[j];
      const float2 mul = ln_w != nullptr ? mod_mul[j] : make_float2(__fadd_rn(1.0f, mod_mul[j].x), __fadd_rn(1.0f, mod_mul[j].y));
#pragma unroll
      for (int p = 0; p < DW_PIX; ++p) {
        if (!INTERIOR && (h0 + 2 * wrow + p / DW_TW >= H || w0 + p % DW_TW >= W)) continue;      // warp-uniform
        // ((x - mean) * rstd) * mul + add, one rounding per operation as before, two channels per instruction
        float2 v = add2(mul2(mul2(add2(res[j][p], splat2(-mean[p])), splat2(rstd[p])), mul), add);
        const int64_t o = o00 + ((int64_t)(p / DW_TW) * W + p % DW_TW) * C + c;
        if constexpr (MODE == 0) {
          *reinterpret_cast<float2*>(y + o) = v;
        } else {
          // A operand of the tensor-core fc1 GEMM: p0 = rn16(v), p1 = rn16(v - p0), p2 = rn16(v - p0 - p1)
          constexpr bool F16 = MODE >= 4;
          constexpr int NP = MODE == 4 ? 2 : (MODE == 5 ? 1 : MODE);
          if constexpr (NP == 1) {
            *reinterpret_cast<uint32_t*>(y0 + o) = pack2<F16>(v.x, v.y);
          } else {
            *reinterpret_cast<uint32_t*>(y0 + o) = split_next<F16>(v);
            if constexpr (NP == 2) {
              *reinterpret_cast<uint32_t*>(y1 + o) = pack2<F16>(v.x, v.y);
            } else {
              *reinterpret_cast<uint32_t*>(y1 + o) = split_next<F16>(v);
              *reinterpret_cast<uint32_t*>(y2 + o) = pack2<F16>(v.x, v.y);
            }
          }
        }
      }
    }
  };
  auto emit_mode = [&](auto interior_c) {
    using std::integral_constant;
    if (y != nullptr) emit(integral_constant<int, 0>{}, interior_c);
    else if (f16 && y1 == nullptr) emit(integral_constant<int, 5>{}, interior_c);
    else if (f16) emit(integral_constant<int, 4>{}, interior_c);
    else if (y2 != nullptr) emit(integral_constant<int, 3>{}, interior_c);
    else if (y1 != nullptr) emit(integral_constant<int, 2>{}, interior_c);
    else emit(integral_constant<int, 1>{}, interior_c);
  };
  if (interior) emit_mode(std::true_type{}); else emit_mode(std::false_type{});
  // No exit barrier: the only remote accesses are the peers' st.async into ln_recv, and this CTA passed its own
  // mbarrier wait only after every one of them had landed.
}

typedef CUresult (*DwEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static DwEncodeTiledFn dw_encode_fn() {
  static DwEncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<DwEncodeTiledFn>(ptr);
  });
  return fn;
}

template <int NJ, int KS, int CL>
static int launch_dwln(const float* x, const float* dw_w, const float* dw_b, const float* ada,
                       int64_t ada_stride, int64_t ada_off, const float* ln_w, const float* ln_b,
                       float* y, __nv_bfloat16* y0, __nv_bfloat16* y1, __nv_bfloat16* y2, int f16,
                       int B, int H, int W, cudaStream_t stream) {
  constexpr int HW_ = DW_TW + KS - 1, HH_ = DW_TH + KS - 1, C = NJ * DW_CH * CL;
  constexpr int smem = (NJ > 1 ? 2 : 1) * HH_ * HW_ * DW_CH * 4;
  static bool configured = false;
  if (!configured) {
    LVAE_CUDA_CALL(cudaFuncSetAttribute(dwln_kernel<NJ, KS, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  DwEncodeTiledFn enc = dw_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return LVAE_E_UNSUPPORTED; }
  // NHWC fp32 viewed as a 4-D tensor (C, W, H, B); box = (64 channels, 8 + k - 1, 16 + k - 1, 1); out-of-image -> 0
  CUtensorMap map;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  cuuint32_t box[4] = {DW_CH, HW_, HH_, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (dwconv input) failed: %d", (int)r); return LVAE_E_BADARG; }
  const int tiles_x = (W + DW_TW - 1) / DW_TW, tiles_y = (H + DW_TH - 1) / DW_TH;
  const int64_t blocks = (int64_t)B * tiles_x * tiles_y;
  // optional L2 prefetch of the tile a later CTA will want, in waves of resident CTAs (4 per SM: 148 * 4 / CL tiles).
  // Measured (profiles/r2_dwln.md): no effect at 1 or 2 waves -- the input was just written by the previous kernel and
  // is largely L2-resident already -- so it is off unless LVAE_DW_L2_AHEAD says otherwise.
  static int l2_ahead = -1;
  if (l2_ahead < 0) { const char* e = getenv("LVAE_DW_L2_AHEAD"); l2_ahead = e ? atoi(e) : 0; }
  const int ahead_tiles = l2_ahead > 0 ? l2_ahead * (148 * 4 / CL) : 0;
  if (CL == 1) {
    LVAE_CUDA_CALL(launch_pdl(dwln_kernel<NJ, KS, CL>, dim3((unsigned)blocks), dim3(32 * DW_WARPS), (size_t)smem, stream,
        map, dw_w, dw_b, ada, ada_stride, ada_off, ln_w, ln_b, y, y0, y1, y2, f16, H, W, tiles_x, tiles_y, B, ahead_tiles));
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(blocks * CL)); cfg.blockDim = dim3(32 * DW_WARPS); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
    LVAE_CUDA_CALL(cudaLaunchKernelEx(&cfg, dwln_kernel<NJ, KS, CL>, map, dw_w, dw_b, ada, ada_stride, ada_off, ln_w, ln_b,
                                      y, y0, y1, y2, f16, H, W, tiles_x, tiles_y, B, ahead_tiles));
  }
  LVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

template <int NJ, int CL>
static int dispatch_k(int k, const float* x, const float* dw_w, const float* dw_b, const float* ada,
                      int64_t ada_stride, int64_t ada_off, const float* ln_w, const float* ln_b,
                      float* y, __nv_bfloat16* y0, __nv_bfloat16* y1, __nv_bfloat16* y2, int f16,
                      int B, int H, int W, cudaStream_t stream) {
  switch (k) {
    case 1: return launch_dwln<NJ, 1, CL>(x, dw_w, dw_b, ada, ada_stride, ada_off, ln_w, ln_b, y, y0, y1, y2, f16, B, H, W, stream);
    case 3: return launch_dwln<NJ, 3, CL>(x, dw_w, dw_b, ada, ada_stride, ada_off, ln_w, ln_b, y, y0, y1, y2, f16, B, H, W, stream);
    case 5: return launch_dwln<NJ, 5, CL>(x, dw_w, dw_b, ada, ada_stride, ada_off, ln_w, ln_b, y, y0, y1, y2, f16, B, H, W, stream);
    case 7: return launch_dwln<NJ, 7, CL>(x, dw_w, dw_b, ada, ada_stride, ada_off, ln_w, ln_b, y, y0, y1, y2, f16, B, H, W, stream);
    default: set_error("dwconv kernel size %d unsupported", k); return LVAE_E_UNSUPPORTED;
  }
}

}  // namespace lvae

static int dwln_dispatch(const float* x, const float* dw_w, const float* dw_b,
                         const float* ada, int64_t ada_stride, int64_t ada_off,
                         const float* ln_w, const float* ln_b, float* y, void* y0, void* y1, void* y2, int f16,
                         int B, int H, int W, int C, int k, void* stream) {
  using namespace lvae;
  LVAE_CHECK_ARG(x && dw_w && dw_b && (y || y0) && (ada || ln_w));
  LVAE_CHECK_ARG(B > 0 && H > 0 && W > 0 && C > 0 && C % 64 == 0);
  LVAE_CHECK_ARG(ln_w != nullptr || (ada_off % 2 == 0 && ada_stride % 2 == 0));      // float2 loads of shift / scale
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* p0 = (__nv_bfloat16*)y0; __nv_bfloat16* p1 = (__nv_bfloat16*)y1; __nv_bfloat16* p2 = (__nv_bfloat16*)y2;
#define LVAE_DWLN_CASE(n64, nj, cl) case n64: return dispatch_k<nj, cl>(k, x, dw_w, dw_b, ada, ada_stride, ada_off, ln_w, ln_b, y, p0, p1, p2, f16, B, H, W, st);
  // C = 64 * NJ * CL: one 64-channel chunk per CTA (everything stays in registers, two CTAs per SM), the channel
  // dimension spread over a thread-block cluster; the rd model's 640 / 768 use two chunks per CTA (cluster <= 8)
  switch (C / 64) {
    LVAE_DWLN_CASE(1, 1, 1) LVAE_DWLN_CASE(2, 1, 2) LVAE_DWLN_CASE(3, 1, 3) LVAE_DWLN_CASE(4, 1, 4)
    LVAE_DWLN_CASE(6, 1, 6) LVAE_DWLN_CASE(8, 1, 8) LVAE_DWLN_CASE(10, 2, 5) LVAE_DWLN_CASE(12, 2, 6)
    default: set_error("dwconv channel count %d unsupported (need C/64 in {1,2,3,4,6,8,10,12})", C); return LVAE_E_UNSUPPORTED;
  }
#undef LVAE_DWLN_CASE
}

extern "C" int lvae_dwconv_ln_adaln(const float* x, const float* dw_w, const float* dw_b,
                                    const float* ada, int64_t ada_stride, int64_t ada_off,
                                    const float* ln_w, const float* ln_b,
                                    float* y, int B, int H, int W, int C, int k, void* stream) {
  LVAE_CHECK_ARG(y != nullptr);
  return dwln_dispatch(x, dw_w, dw_b, ada, ada_stride, ada_off, ln_w, ln_b, y, nullptr, nullptr, nullptr, 0, B, H, W, C, k, stream);
}

extern "C" int lvae_dwconv_ln_adaln_planes(const float* x, const float* dw_w, const float* dw_b,
                                           const float* ada, int64_t ada_stride, int64_t ada_off,
                                           const float* ln_w, const float* ln_b,
                                           void* y0, void* y1, void* y2, int plane_format,
                                           int B, int H, int W, int C, int k, void* stream) {
  LVAE_CHECK_ARG(y0 != nullptr && (y2 == nullptr || y1 != nullptr));
  LVAE_CHECK_ARG(plane_format == LVAE_PLANES_BF16 || plane_format == LVAE_PLANES_F16);
  LVAE_CHECK_ARG(plane_format != LVAE_PLANES_F16 || y2 == nullptr);    // fp16 operands travel as 2 planes (F16X3) or 1 (F16)
  return dwln_dispatch(x, dw_w, dw_b, ada, ada_stride, ada_off, ln_w, ln_b, nullptr, y0, y1, y2,
                       plane_format == LVAE_PLANES_F16, B, H, W, C, k, stream);
}
