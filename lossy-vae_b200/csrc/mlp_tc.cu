// Fused ConvNeXt MLP for the narrow, position-rich layers (C <= 192: the H/4 stages of qarv_base and qres34m):
//
//   out[m, :] = res[m, :] + gamma * ( W2 gelu(W1 a[m, :] + b1) + b2 )          (lvae/models/common.py:154-160)
//
// in ONE kernel, so the hidden activation [M, hidden] never travels through HBM: the unfused pair (fc1 writes two
// 16-bit planes of the hidden tensor, fc2 reads them back together with the residual) moves 28*C bytes per position
// at mlp_ratio 2, this kernel 8*C + the L2-resident weights.  Arithmetic, operand planes, MMA order per output
// element and the GELU code are those of gemm_tc_kernel<2, *>, so the result is bit-identical to the unfused path
// (tests/test_gpu_kernels.py::test_fused_mlp_equals_unfused).
//
// One persistent CTA per SM, 128 rows per tile, the hidden dimension in chunks of 32:
//   warp 0   TMA producer: the [128 x C] A tile (both planes) once per tile and per chunk the [32 x C] slice of W1
//   warp 3   TMA producer of the [C x 32] slices of W2 (independent 2-stage rings)
//   warp 1   fc1 issuer: acc1[j & 1] = A W1_j^T (N = 32), up to two chunks ahead of the GELU warps
//   warp 2   fc2 issuer: acc2 += H_j W2_j^T (N = C) as soon as the GELU warps have published chunk j
//   warps 4-19  GELU epilogue, two groups of 8 warps alternating chunks: tcgen05.ld 32 rows x 16 columns of acc1
//            (main + cross), bias + GELU, split into two
//            16-bit planes and store them straight into the SWIZZLE_64B K-major shared-memory tile that fc2 reads
//            as its A operand (fence.proxy.async + mbarrier hand-over); then together the tile epilogue (layer scale +
//            residual, 128-bit rows staged through the idle hidden buffers)
// acc1 (TMEM) and the hidden chunk (shared memory) are double-buffered: fc1 runs one chunk ahead of the GELU warps and
// fc2 one chunk behind, so neither side waits for the other's latency chain (single-buffered, the ld -> release ->
// MMA -> commit -> wake round trip made every chunk 5 k cycles long).
// TMEM (512 columns): acc2 main | acc2 cross (2 x C) | 2 x (acc1 main | acc1 cross) (2 x 2 x 32).
#include "tc_common.cuh"
#include <cuda_bf16.h>
#include <stdlib.h>

namespace lvae {

constexpr int ML_BM = 128;
constexpr int ML_HB = 32;                         // hidden units per chunk (= K of one fc2 step, 64-byte operand rows)
constexpr int ML_EPI_WARPS = 16;
constexpr int ML_THREADS = 128 + 32 * ML_EPI_WARPS;     // A/W1 producer, fc1 issuer, fc2 issuer, W2 producer, GELU / epilogue warps

struct MlpParams {
  int M, C, HID, num_tiles;
  const float* b1; const float* b2; const float* gamma; const float* res; float* out;
  uint16_t* out_pl[2]; int pl_act;       // optional planes of the result (pl_act: of gelu(result))
  float acc_scale; int f16;
};
struct MlpMaps { CUtensorMap a[2]; CUtensorMap w1[2]; CUtensorMap w2[2]; };

__device__ __forceinline__ void tc_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr) : "memory");
}

template <int NKB>
__global__ void __launch_bounds__(ML_THREADS, 1)
mlp_tc_kernel(const __grid_constant__ MlpMaps maps, const MlpParams p) {
  extern __shared__ __align__(1024) uint8_t ml_smem_raw[];
  uint8_t* smem = ml_smem_raw + ((1024u - (smem_u32(ml_smem_raw) & 1023u)) & 1023u);
  constexpr int C = NKB * 64;
  constexpr int NST = NKB == 3 ? 2 : 4;                 // weight-ring depth: whatever shared memory is left
  const int NCH = p.HID / ML_HB;
  // shared-memory map
  const int a_tile = ML_BM * 64 * 2;                    // one plane of one 64-wide k-block of A: 16 KB
  const int a_bytes = NKB * 2 * a_tile;
  const int w1_tile = ML_HB * 64 * 2;                   // [32 x 64] slice of a W1 plane: 4 KB
  const int w2_tile = C * ML_HB * 2;                    // [C x 32] slice of a W2 plane (64-byte rows)
  const int w1_stage = NKB * 2 * w1_tile, w2_stage = 2 * w2_tile;   // separate 2-stage rings: a W1 slice is released as
  const int h_tile = ML_BM * ML_HB * 2;                 // soon as fc1 has read it, a W2 slice only after fc2 (much later)
  uint8_t* a_s = smem;
  uint8_t* w1_s = a_s + a_bytes;
  uint8_t* w2_s = w1_s + NST * w1_stage;
  uint8_t* h_s = w2_s + NST * w2_stage;                   // two hidden-chunk buffers of 2 planes each (32 KB); between
  const int h_buf = 2 * h_tile;                         // tiles the same 32 KB stage the tile epilogue (16 x 2 KB)
  uint64_t* bars = reinterpret_cast<uint64_t*>(h_s + 2 * h_buf);
  uint64_t* a_full = bars;        uint64_t* a_empty = bars + 1;
  uint64_t* w1_full = bars + 2;   uint64_t* w1_empty = bars + 6;       // [NST <= 4] each
  uint64_t* w2_full = bars + 10;  uint64_t* w2_empty = bars + 14;      // [NST <= 4] each
  uint64_t* acc1_full = bars + 18; uint64_t* acc1_free = bars + 20;    // [2] each: acc1 is double-buffered in TMEM
  uint64_t* h_ready = bars + 22;  uint64_t* h_free = bars + 24;        // [2] each: so is the hidden chunk in smem
  uint64_t* acc2_full = bars + 26; uint64_t* acc2_free = bars + 27;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 28);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    mbar_init(smem_u32(a_full), 1); mbar_init(smem_u32(a_empty), 1);
    for (int s = 0; s < NST; ++s) {
      mbar_init(smem_u32(w1_full + s), 1); mbar_init(smem_u32(w1_empty + s), 1);
      mbar_init(smem_u32(w2_full + s), 1); mbar_init(smem_u32(w2_empty + s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(acc1_full + s), 1); mbar_init(smem_u32(acc1_free + s), ML_EPI_WARPS / 2);
      mbar_init(smem_u32(h_ready + s), ML_EPI_WARPS / 2); mbar_init(smem_u32(h_free + s), 1);
    }
    mbar_init(smem_u32(acc2_full), 1); mbar_init(smem_u32(acc2_free), ML_EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a[0]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.w1[0]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.w2[0]) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();                                           // PDL: the set-up above overlapped the predecessor's tail
  const uint32_t t_acc2 = tmem_base, t_acc2x = tmem_base + (uint32_t)C;
  const uint32_t t_acc1 = tmem_base + (uint32_t)(2 * C);          // buffer b: main at + b * 64, cross at + b * 64 + 32

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      auto load_w1 = [&](uint32_t gg) {
        const int s = (int)(gg % NST), j = (int)(gg % (uint32_t)NCH);
        mbar_wait(smem_u32(w1_empty + s), ((gg / NST) & 1) ^ 1);
        const uint32_t wf1 = smem_u32(w1_full + s);
        mbar_expect_tx(wf1, (uint32_t)w1_stage);
        for (int kb = 0; kb < NKB; ++kb)
          for (int pl = 0; pl < 2; ++pl)
            tma_load_2d(smem_u32(w1_s + s * w1_stage + (kb * 2 + pl) * w1_tile), &maps.w1[pl], wf1, kb * 64, j * ML_HB);
      };
      int it = 0; uint32_t g = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
        mbar_wait(smem_u32(a_empty), ((uint32_t)it & 1) ^ 1);
        const uint32_t af = smem_u32(a_full);
        mbar_expect_tx(af, (uint32_t)a_bytes);
        for (int kb = 0; kb < NKB; ++kb)
          for (int pl = 0; pl < 2; ++pl)
            tma_load_2d(smem_u32(a_s + (kb * 2 + pl) * a_tile), &maps.a[pl], af, kb * 64, t * ML_BM);
        // next tile's A operand: pull it into L2 now, so that the load issued when this tile's last fc1 retires is short
        if (t + (int)gridDim.x < p.num_tiles)
          for (int kb = 0; kb < NKB; ++kb)
            for (int pl = 0; pl < 2; ++pl) tma_prefetch_2d(&maps.a[pl], kb * 64, (t + (int)gridDim.x) * ML_BM);
        for (int j = 0; j < NCH; ++j, ++g) load_w1(g);
      }
    }
  } else if (warp == 3) {
    // ============================ W2 producer (own warp: its stage frees only after the late fc2 of a chunk, and the
    //                              W1 / A stream must not queue behind that wait) ============================
    if (lane == 0) {
      uint32_t g = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x)
        for (int j = 0; j < NCH; ++j, ++g) {
          const int s = (int)(g % NST);
          mbar_wait(smem_u32(w2_empty + s), ((g / NST) & 1) ^ 1);
          const uint32_t wf2 = smem_u32(w2_full + s);
          mbar_expect_tx(wf2, (uint32_t)w2_stage);
          for (int pl = 0; pl < 2; ++pl)
            tma_load_2d(smem_u32(w2_s + s * w2_stage + pl * w2_tile), &maps.w2[pl], wf2, j * ML_HB, 0);
        }
    }
  } else if (warp == 1 || warp == 2) {
    // ============================ MMA issuers ============================
    const uint32_t fmt = p.f16 ? 0u : 1u;
    const uint32_t idesc1 = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(ML_HB >> 3) << 17) | ((uint32_t)(ML_BM >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(ML_BM >> 4) << 24);
    // N doubled: the two weight planes are adjacent in a stage and main | cross adjacent in TMEM (see gemm_tc.cu)
    const uint32_t idesc1n = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(ML_HB >> 2) << 17) | ((uint32_t)(ML_BM >> 4) << 24);
    const uint32_t idesc2n = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(C >> 2) << 17) | ((uint32_t)(ML_BM >> 4) << 24);
    // Every shared-memory descriptor is loop-invariant (two stage buffers): build them once.  The MMAs here are short
    // (N = 32 / K = 32), so the single issuing thread is the critical resource: ~4 instructions per MMA, not ~15.
    uint64_t dA[NKB][2], dW1[NKB][2], dW2[2], dH[2][2];       // W descriptors of stage 0; stage s adds (s * stage bytes) >> 4
#pragma unroll
    for (int kb = 0; kb < NKB; ++kb)
#pragma unroll
      for (int pl = 0; pl < 2; ++pl) {
        dA[kb][pl] = make_desc(smem_u32(a_s + (kb * 2 + pl) * a_tile), 64);
        dW1[kb][pl] = make_desc(smem_u32(w1_s + (kb * 2 + pl) * w1_tile), 64);
      }
#pragma unroll
    for (int pl = 0; pl < 2; ++pl) {
      dW2[pl] = make_desc(smem_u32(w2_s + pl * w2_tile), ML_HB);
#pragma unroll
      for (int b = 0; b < 2; ++b) dH[b][pl] = make_desc(smem_u32(h_s + b * h_buf + pl * h_tile), ML_HB);
    }
    int it = 0; uint32_t g = 0;
    // Two issuing warps: if one thread issued both GEMMs it would sit in the wait for the GELU warps' hidden chunk
    // (fc2) while the next fc1 -- whose accumulator buffer is long free -- is not even issued, and every chunk would
    // pay the full MMA -> commit -> wake latency chain.  tcgen05.commit tracks the MMAs of the issuing thread only.
    if (warp == 1) {
      auto issue_fc1 = [&](uint32_t gg, bool last_of_tile) {
        // runs up to two chunks ahead of the GELU warps: its accumulator buffer was released when they loaded chunk gg - 2
        const int b = gg & 1; const uint32_t ph = (gg >> 1) & 1;
        const int ws = (int)(gg % NST);
        mbar_wait(smem_u32(w1_full + ws), (gg / NST) & 1);
        mbar_wait(smem_u32(acc1_free + b), ph ^ 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t d_main = t_acc1 + (uint32_t)(b * 2 * ML_HB), d_cross = d_main + ML_HB;
#pragma unroll
          for (int kb = 0; kb < NKB; ++kb) {
            const uint64_t da0 = dA[kb][0], da1 = dA[kb][1];
            const uint64_t wo = (uint64_t)((ws * w1_stage) >> 4);
            const uint64_t db0 = dW1[kb][0] + wo, db1 = dW1[kb][1] + wo;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t ko = (uint64_t)(k * 2);
              tc_mma(d_main, da0 + ko, db0 + ko, idesc1n, (kb | k) ? 1u : 0u);       // a0 * [w0; w1] -> main | cross
              tc_mma(d_cross, da1 + ko, db0 + ko, idesc1, 1u);
            }
          }
          tc_commit(smem_u32(acc1_full + b));
          tc_commit(smem_u32(w1_empty + ws));
          if (last_of_tile) tc_commit(smem_u32(a_empty));          // the A tile may be overwritten once these MMAs are done
        }
        __syncwarp();
      };
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
        mbar_wait(smem_u32(a_full), (uint32_t)it & 1);
        for (int j = 0; j < NCH; ++j, ++g) issue_fc1(g, j + 1 == NCH);
      }
    } else {
      auto issue_fc2 = [&](uint32_t gp, int jp, int itp) {          // chunk gp (index jp within its tile itp)
        const int b = gp & 1; const uint32_t ph = (gp >> 1) & 1;
        const int ws = (int)(gp % NST);
        mbar_wait(smem_u32(w2_full + ws), (gp / NST) & 1);
        mbar_wait(smem_u32(h_ready + b), ph);
        if (jp == 0) mbar_wait(smem_u32(acc2_free), ((uint32_t)itp & 1) ^ 1);   // previous tile's result has been read
        tc_fence_after();
        if (lane == 0) {
          const uint64_t ha0 = b ? dH[1][0] : dH[0][0], ha1 = b ? dH[1][1] : dH[0][1];
          const uint64_t wo = (uint64_t)((ws * w2_stage) >> 4);
          const uint64_t wb0 = dW2[0] + wo, wb1 = dW2[1] + wo;
#pragma unroll
          for (int k = 0; k < ML_HB / 16; ++k) {
            const uint64_t ko = (uint64_t)(k * 2);
            const uint32_t first = (jp | k) ? 1u : 0u;
            if (2 * C <= 256) {
              tc_mma(t_acc2, ha0 + ko, wb0 + ko, idesc2n, first);                    // h0 * [w0; w1] -> main | cross
              tc_mma(t_acc2x, ha1 + ko, wb0 + ko, idesc2, 1u);
            } else {                                                                 // N = 2 C would exceed 256
              tc_mma(t_acc2, ha0 + ko, wb0 + ko, idesc2, first);
              tc_mma(t_acc2x, ha0 + ko, wb1 + ko, idesc2, first);
              tc_mma(t_acc2x, ha1 + ko, wb0 + ko, idesc2, 1u);
            }
          }
          tc_commit(smem_u32(h_free + b));
          tc_commit(smem_u32(w2_empty + ws));
          if (jp == NCH - 1) tc_commit(smem_u32(acc2_full));
        }
        __syncwarp();
      };
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it)
        for (int j = 0; j < NCH; ++j, ++g) issue_fc2(g, j, it);
    }
  } else {
    // ============================ epilogue warps ============================
    const int ew = warp - 4;
    const int q = warp & 3;                                  // TMEM lane quarter this warp may access
    // Two groups of 8 warps alternate chunks (group = accumulator / hidden buffer index): one chunk's chain of
    // latencies (barrier wake, TMEM load, GELU, proxy fence, hand-over to fc2) overlaps the other group's.
    const int grp = ew >> 3;
    const int sub = (ew & 7) >> 2;                           // 16-column half of the 32-wide hidden chunk
    const int psub = ew >> 2;                                // tile epilogue: 16-column pieces psub, psub + 4, ...
    const bool f16 = p.f16 != 0;
    const int row = q * 32 + lane;
    const uint32_t hsw = (uint32_t)((row >> 1) & 3);         // SWIZZLE_64B: 16-byte unit ^= (addr >> 7) & 3
    const uint32_t h_off0 = (uint32_t)(row * 64) + (((uint32_t)(2 * sub) ^ hsw) << 4);
    const uint32_t h_off1 = (uint32_t)(row * 64) + (((uint32_t)(2 * sub + 1) ^ hsw) << 4);
    uint32_t* stg = reinterpret_cast<uint32_t*>(h_s) + ew * 512;                       // 2 KB of the (idle) hidden buffers
    const float2 sc = splat2(p.acc_scale);
    const int b = grp;                                       // this group's acc1 / hidden buffer
    uint8_t* hb = h_s + b * h_buf;
    int it = 0; uint32_t g = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      const int m0 = t * ML_BM;
      for (int j = 0; j < NCH; ++j, ++g) {
        if ((int)(g & 1) != grp) continue;
        const uint32_t ph = (g >> 1) & 1;
        const int nb = j * ML_HB + sub * 16;
        float2 bias[8];                                      // in flight while this warp waits for the accumulator
#pragma unroll
        for (int i = 0; i < 8; ++i) bias[i] = __ldg(reinterpret_cast<const float2*>(p.b1 + nb + 2 * i));
        mbar_wait(smem_u32(acc1_full + b), ph);
        tc_fence_after();
        uint32_t vm[32], vx[32];
        const uint32_t ta = t_acc1 + (uint32_t)(b * 2 * ML_HB) + ((uint32_t)(q * 32) << 16) + (uint32_t)(sub * 16);
        tc_ld16(ta, vm);
        tc_ld16(ta + ML_HB, vx);
        tc_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(acc1_free + b));
        uint32_t w0[8], w1[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2 sum = mul2(add2(make_float2(__uint_as_float(vm[2 * i]), __uint_as_float(vm[2 * i + 1])),
                                       make_float2(__uint_as_float(vx[2 * i]), __uint_as_float(vx[2 * i + 1]))), sc);
          float2 gv = gelu_erf2(add2(sum, bias[i]));
          w0[i] = split_next(gv, f16);
          w1[i] = f16 ? pack2<true>(gv.x, gv.y) : pack2<false>(gv.x, gv.y);
        }
        mbar_wait(smem_u32(h_free + b), ph ^ 1);             // fc2 of chunk g - 2 has consumed this buffer
        *reinterpret_cast<uint4*>(hb + h_off0) = make_uint4(w0[0], w0[1], w0[2], w0[3]);
        *reinterpret_cast<uint4*>(hb + h_off1) = make_uint4(w0[4], w0[5], w0[6], w0[7]);
        *reinterpret_cast<uint4*>(hb + h_tile + h_off0) = make_uint4(w1[0], w1[1], w1[2], w1[3]);
        *reinterpret_cast<uint4*>(hb + h_tile + h_off1) = make_uint4(w1[4], w1[5], w1[6], w1[7]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA's async proxy
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(h_ready + b));
      }
      // ---- tile epilogue, all 16 warps: out = res + gamma * (acc2 + b2).  Warp (q, sub) takes the 16-column pieces
      //      sub, sub + 4, ... of its lane quarter, staged through its 2 KB slice of the hidden buffers (idle: every
      //      fc2 of this tile has completed, and nobody starts the next tile's chunks before the bar.sync below)
      mbar_wait(smem_u32(acc2_full), (uint32_t)it & 1);
      tc_fence_after();
      const int npieces = C / 16;
      int last_piece = psub; while (last_piece + 4 < npieces) last_piece += 4;
      for (int c = psub; c < npieces; c += 4) {
        uint32_t v[32], u[32];
        const uint32_t ta = ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 16);
        tc_ld16(t_acc2 + ta, v);
        tc_ld16(t_acc2x + ta, u);
        tc_wait_ld();
        if (c == last_piece) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(acc2_free));
        }
#pragma unroll
        for (int jj = 0; jj < 16; jj += 2) {
          const float2 sum = mul2(add2(make_float2(__uint_as_float(v[jj]), __uint_as_float(v[jj + 1])),
                                       make_float2(__uint_as_float(u[jj]), __uint_as_float(u[jj + 1]))), sc);
          v[jj] = __float_as_uint(sum.x); v[jj + 1] = __float_as_uint(sum.y);
        }
        // 32 rows x 64 bytes, 16-byte units XOR-swizzled by (row >> 1) & 3: conflict-free row writes and 4-lane row reads
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4)
          *reinterpret_cast<uint4*>(stg + lane * 16 + ((j4 ^ ((lane >> 1) & 3)) << 2)) = make_uint4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
        __syncwarp();
        const int srow = lane >> 2, l4 = lane & 3;           // 8 rows per pass, 4 lanes x 4 columns per row
        const int n = c * 16 + 4 * l4;
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.b2 + n));
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.gamma + n));
        const int row0 = m0 + q * 32;
        const int64_t o0 = (int64_t)(row0 + srow) * C + n;
        const int rows = p.M - row0 - srow;                  // valid while 8 * i < rows
        float4 rr[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          rr[i] = (8 * i < rows) ? *reinterpret_cast<const float4*>(p.res + o0 + (int64_t)(8 * i) * C) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (8 * i >= rows) break;
          const int r = 8 * i + srow;
          const float4 a = *reinterpret_cast<const float4*>(stg + r * 16 + ((l4 ^ ((r >> 1) & 3)) << 2));
          float4 x;
          x.x = __fadd_rn(__fmul_rn(__fadd_rn(a.x, b4.x), g4.x), rr[i].x); x.y = __fadd_rn(__fmul_rn(__fadd_rn(a.y, b4.y), g4.y), rr[i].y);
          x.z = __fadd_rn(__fmul_rn(__fadd_rn(a.z, b4.z), g4.z), rr[i].z); x.w = __fadd_rn(__fmul_rn(__fadd_rn(a.w, b4.w), g4.w), rr[i].w);
          *reinterpret_cast<float4*>(p.out + o0 + (int64_t)(8 * i) * C) = x;
          if (p.out_pl[0] != nullptr) {
            if (p.pl_act) { x.x = gelu_erf(x.x); x.y = gelu_erf(x.y); x.z = gelu_erf(x.z); x.w = gelu_erf(x.w); }
            float2 lo = make_float2(x.x, x.y), hi = make_float2(x.z, x.w);
            uint2 w;
            w.x = split_next(lo, f16); w.y = split_next(hi, f16);
            *reinterpret_cast<uint2*>(p.out_pl[0] + o0 + (int64_t)(8 * i) * C) = w;
            w.x = f16 ? pack2<true>(lo.x, lo.y) : pack2<false>(lo.x, lo.y);
            w.y = f16 ? pack2<true>(hi.x, hi.y) : pack2<false>(hi.x, hi.y);
            *reinterpret_cast<uint2*>(p.out_pl[1] + o0 + (int64_t)(8 * i) * C) = w;
          }
        }
        __syncwarp();
      }
      // the hidden buffers go back to their day job
      asm volatile("bar.sync 1, %0;" ::"n"(32 * ML_EPI_WARPS) : "memory");
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

}  // namespace lvae

using namespace lvae;

extern "C" int lvae_convnext_mlp(const void* a_p0, const void* a_p1, const void* w1_p0, const void* w1_p1, const float* b1,
                                 const void* w2_p0, const void* w2_p1, const float* b2, const float* gamma,
                                 const float* res, float* out, int64_t M, int C, int hidden, int precision, void* stream) {
  return lvae_convnext_mlp_planes(a_p0, a_p1, w1_p0, w1_p1, b1, w2_p0, w2_p1, b2, gamma, res, out, nullptr, nullptr, 0,
                                  M, C, hidden, precision, stream);
}

extern "C" int lvae_convnext_mlp_planes(const void* a_p0, const void* a_p1, const void* w1_p0, const void* w1_p1, const float* b1,
                                        const void* w2_p0, const void* w2_p1, const float* b2, const float* gamma,
                                        const float* res, float* out, void* out_p0, void* out_p1, int out_planes_gelu,
                                        int64_t M, int C, int hidden, int precision, void* stream) {
  LVAE_CHECK_ARG(a_p0 && a_p1 && w1_p0 && w1_p1 && w2_p0 && w2_p1 && b1 && b2 && gamma && res && out);
  LVAE_CHECK_ARG((out_p0 == nullptr) == (out_p1 == nullptr));
  LVAE_CHECK_ARG(M > 0 && M < (1ll << 31));
  LVAE_CHECK_ARG(C % 64 == 0 && C >= 64 && C <= 192 && hidden % ML_HB == 0 && hidden >= ML_HB);
  LVAE_CHECK_ARG(precision == LVAE_PREC_F16X3 || precision == LVAE_PREC_BF16X3);
  MlpParams p;
  p.M = (int)M; p.C = C; p.HID = hidden; p.num_tiles = (int)((M + ML_BM - 1) / ML_BM);
  p.b1 = b1; p.b2 = b2; p.gamma = gamma; p.res = res; p.out = out;
  p.out_pl[0] = (uint16_t*)out_p0; p.out_pl[1] = (uint16_t*)out_p1; p.pl_act = out_planes_gelu;
  p.f16 = precision == LVAE_PREC_F16X3 ? 1 : 0;
  p.acc_scale = p.f16 ? 1.0f / LVAE_F16_WEIGHT_SCALE : 1.0f;
  MlpMaps maps;
  const void* ap[2] = {a_p0, a_p1}; const void* w1p[2] = {w1_p0, w1_p1}; const void* w2p[2] = {w2_p0, w2_p1};
  int rc;
  for (int i = 0; i < 2; ++i) {
    if ((rc = make_map(&maps.a[i], ap[i], M, C, ML_BM, 64))) return rc;
    if ((rc = make_map(&maps.w1[i], w1p[i], hidden, C, ML_HB, 64))) return rc;
    if ((rc = make_map(&maps.w2[i], w2p[i], C, hidden, C, ML_HB))) return rc;
  }
  const int NKB = C / 64;
  const int nst = NKB == 3 ? 2 : 4;
  const int smem = 1024 + NKB * 2 * (ML_BM * 64 * 2) + nst * (NKB * 2 * (ML_HB * 64 * 2) + 2 * (C * ML_HB * 2)) +
                   2 * 2 * (ML_BM * ML_HB * 2) + 256;
  LVAE_CHECK_ARG(smem <= 227 * 1024);
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    LVAE_CUDA_CALL(cudaFuncSetAttribute(mlp_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    LVAE_CUDA_CALL(cudaFuncSetAttribute(mlp_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    LVAE_CUDA_CALL(cudaFuncSetAttribute(mlp_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  const int grid = p.num_tiles < n_sm ? p.num_tiles : n_sm;
  if (NKB == 3) LVAE_CUDA_CALL(launch_pdl(mlp_tc_kernel<3>, dim3(grid), dim3(ML_THREADS), (size_t)smem, (cudaStream_t)stream, maps, p));
  else if (NKB == 2) LVAE_CUDA_CALL(launch_pdl(mlp_tc_kernel<2>, dim3(grid), dim3(ML_THREADS), (size_t)smem, (cudaStream_t)stream, maps, p));
  else LVAE_CUDA_CALL(launch_pdl(mlp_tc_kernel<1>, dim3(grid), dim3(ML_THREADS), (size_t)smem, (cudaStream_t)stream, maps, p));
  LVAE_CUDA_LAUNCH_CHECK();
  return 0;
}
