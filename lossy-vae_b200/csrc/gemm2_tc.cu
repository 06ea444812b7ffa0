// cta_group::2 tensor-core GEMM: the large [M,K] x [N,K]^T contractions of the path on CTA PAIRS (2-plane modes).
//
//   out[m, n] = epilogue( sum_k A[m, k] * W[n, k] + bias[n] )        (same contract as gemm_tc_kernel<2, EK_GELU | EK_ROWS>)
//
// Why: r2 measurements (profiles/r2_mma_probe.md) show that a kind::f16 tcgen05.mma costs N/2 cycles per SM (M = 128 per
// SM) -- NOT a constant -- so the 128 x 128 tiles of gemm_tc_kernel are not issue-bound; they are bound by the bytes an
// SM can pull from L2 (~42 B/clk/SM measured in that kernel): a 128 x 128 tile with two 16-bit planes per operand needs
// 16 KB per k-step for 192 MMA cycles = 85 B/clk.  Here a pair of CTAs (one cluster, two SMs of a TPC) computes a
// 256 x BN tile with ONE instruction stream (tcgen05.mma.cta_group::2, M = 256): each CTA stages its own 128 rows of A
// and only HALF of the W tile (BN/2 rows); the tensor cores of both SMs read both halves.  With BN = 256 that is
// 16 KB per k-step per CTA for 384 MMA cycles = 43 B/clk -- half the operand traffic per flop.
//
// TMEM holds main | cross (see gemm_tc.cu: tcgen05 accumulates with truncation, the correction terms get their own
// accumulator) = 2 BN columns: BN = 256 fills all 512 columns, so there is no second accumulator buffer for BN > 128.
// Instead the 16 epilogue warps DRAIN the whole accumulator into registers first (two 32 x 32 chunks per warp, main +
// cross added once), release TMEM to the MMA warp, and only then run the expensive part (bias / GELU / plane split /
// residual / stores) while the next tile's MMAs are already running.  BN <= 128 double-buffers as before.
//
// Roles per CTA (18 warps): warp 0 TMA producer (both CTAs: own A rows, own half of W; completion on the LEADER's
// mbarrier), warp 1 TMEM allocation + (leader CTA only) MMA issue, tcgen05.commit multicast to both CTAs' barriers,
// warps 2-17 epilogue of the CTA's own 128 rows x BN columns.  K order, operand planes, MMA order per output element
// and epilogue arithmetic are those of gemm_tc_kernel, so results are bit-identical to it
// (tests/test_gpu_kernels.py::test_gemm2_equals_gemm_tc).
#include "tc_common.cuh"
#include <cuda_bf16.h>
#include <stdlib.h>

namespace lvae {

constexpr int G2_BM = 128;                      // rows per CTA; a pair tile is 256 x BN
constexpr int G2_BK = 64;                       // one SWIZZLE_128B row of 16-bit elements
constexpr int G2_EPI_WARPS = 16;
constexpr int G2_THREADS = 64 + 32 * G2_EPI_WARPS;
constexpr int G2_STG = 2048;                    // per-warp transpose buffer: 32 rows x 64 B
constexpr int G2_GELU = 0, G2_ROWS = 1;

struct G2Params {
  int M, N, K, BN, n_tiles, num_pair_tiles, stages, nbuf, k_split;
  const float* bias; const float* gamma; const float* res;
  float* out; uint16_t* out_pl[2];
  int epi, pl_act, f16;
  float acc_scale;
};
struct G2Maps { CUtensorMap a[2]; CUtensorMap b[2]; CUtensorMap a1[2]; };

// cycle breakdown of cluster 0 (diagnostics, lvae_debug_prof(2, ...)): [0] MMA thread total, [1] waiting for operands
// (full), [2] waiting for the accumulator (tempty), [3] producer total, [4] producer waiting for a free stage,
// [5] epilogue warp 2 total, [6] waiting for the accumulator (tfull), [7] draining TMEM, [8] tiles
__device__ unsigned long long g2_prof[16];

__device__ __forceinline__ float g2_gelu_grad(float h) {      // = gelu_grad of gemm_tc.cu
  const float cdf = 0.5f * (1.f + erff(h * 0.70710678118654752f));
  return fmaf(h * 0.39894228040143268f, __expf(-0.5f * h * h), cdf);
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load into this CTA's shared memory, transaction bytes credited to an mbarrier of the pair's leader CTA
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_mma2(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
// arrives (once the issuing thread's MMAs have completed) on the barrier at this shared-memory offset in BOTH CTAs
__device__ __forceinline__ void tc_commit2(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}

// (main + cross) * 2^-s of one 32-row x 32-column accumulator chunk -> s[32]
__device__ __forceinline__ void g2_load_sum(uint32_t taddr, uint32_t cross_off, float scale, uint32_t (&s)[32]) {
  uint32_t u[32];
  tc_ld32(taddr, s);
  tc_ld32(taddr + cross_off, u);
  tc_wait_ld();
  const float2 sc = splat2(scale);
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    const float2 sum = mul2(add2(make_float2(__uint_as_float(s[j]), __uint_as_float(s[j + 1])),
                                 make_float2(__uint_as_float(u[j]), __uint_as_float(u[j + 1]))), sc);
    s[j] = __float_as_uint(sum.x); s[j + 1] = __float_as_uint(sum.y);
  }
}

// fc1 epilogue of one chunk: bias + GELU, split into two 16-bit planes, 64-byte row segments out (gemm_tc.cu EK_GELU)
__device__ __forceinline__ void g2_chunk_gelu(const G2Params& p, uint32_t (&v)[32], uint32_t* stg, int row0, int nb, int lane) {
  const bool f16 = p.f16 != 0;
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    const float2 b2 = p.bias ? __ldg(reinterpret_cast<const float2*>(p.bias + nb + j)) : make_float2(0.f, 0.f);
    const float2 o = gelu_erf2(add2(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), b2));
    v[j] = __float_as_uint(o.x); v[j + 1] = __float_as_uint(o.y);
  }
  const int sub = lane >> 3, l8 = lane & 7;                  // 4 rows per pass, 8 lanes x 4 columns per row
  const int n = nb + 4 * l8;
  const bool full = row0 + 32 <= p.M;
#pragma unroll
  for (int pl = 0; pl < 2; ++pl) {
    if (p.out_pl[pl] == nullptr) break;
#pragma unroll
    for (int j4 = 0; j4 < 4; ++j4) {
      uint32_t w[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = j4 * 8 + e * 2;
        float2 gv = make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1]));
        if (pl == 0) {                                       // leaves the exact residual for the second plane
          w[e] = split_next(gv, f16);
          v[j] = __float_as_uint(gv.x); v[j + 1] = __float_as_uint(gv.y);
        } else {
          w[e] = f16 ? pack2<true>(gv.x, gv.y) : pack2<false>(gv.x, gv.y);
        }
      }
      *reinterpret_cast<uint4*>(stg + lane * 16 + ((j4 ^ ((lane >> 1) & 3)) << 2)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    __syncwarp();
    uint16_t* dst = p.out_pl[pl] + (int64_t)(row0 + sub) * p.N + n;
#pragma unroll
    for (int r = 0; r < 32; r += 4) {
      const uint2 w2 = *reinterpret_cast<const uint2*>(stg + (r + sub) * 16 + ((((l8 >> 1) ^ (((r + sub) >> 1) & 3)) << 2) | ((l8 & 1) << 1)));
      if (full || row0 + r + sub < p.M) *reinterpret_cast<uint2*>(dst + (int64_t)r * p.N) = w2;
    }
    __syncwarp();
  }
}

// bias | layer-scale + residual | bias + residual | GELU' -> fp32 rows (+ planes) of one chunk, in two 16-column pieces
// staged through 32 rows x 64 B (XOR-swizzled 16-byte units: conflict-free row writes and 4-lane row reads)
__device__ __forceinline__ void g2_chunk_rows(const G2Params& p, const uint32_t (&v)[32], uint32_t* stg, int row0, int nb, int lane) {
  const bool f16 = p.f16 != 0;
  const bool scale_res = p.epi == LVAE_EPI_SCALE_RES;
  const bool gelu_bwd = p.epi == LVAE_EPI_GELU_BWD;
  const bool has_res = scale_res || gelu_bwd || p.epi == LVAE_EPI_BIAS_RES;
  const int srow = lane >> 2, l4 = lane & 3;                 // 8 rows per pass, 4 lanes x 4 columns per row
#pragma unroll
  for (int h = 0; h < 2; ++h) {
#pragma unroll
    for (int j4 = 0; j4 < 4; ++j4)
      *reinterpret_cast<uint4*>(stg + lane * 16 + ((j4 ^ ((lane >> 1) & 3)) << 2)) =
          make_uint4(v[16 * h + 4 * j4], v[16 * h + 4 * j4 + 1], v[16 * h + 4 * j4 + 2], v[16 * h + 4 * j4 + 3]);
    __syncwarp();
    const int n = nb + 16 * h + 4 * l4;
    const float4 b4 = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 g4 = scale_res ? __ldg(reinterpret_cast<const float4*>(p.gamma + n)) : make_float4(1.f, 1.f, 1.f, 1.f);
    const int64_t o0 = (int64_t)(row0 + srow) * p.N + n;
    const int rows = p.M - row0 - srow;                      // valid while 8 * i < rows
    float4 rr[4];
    if (has_res) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        rr[i] = (8 * i < rows) ? *reinterpret_cast<const float4*>(p.res + o0 + (int64_t)(8 * i) * p.N) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (8 * i >= rows) break;
      const int r = 8 * i + srow;
      const float4 a = *reinterpret_cast<const float4*>(stg + r * 16 + ((l4 ^ ((r >> 1) & 3)) << 2));
      float4 x = make_float4(__fadd_rn(a.x, b4.x), __fadd_rn(a.y, b4.y), __fadd_rn(a.z, b4.z), __fadd_rn(a.w, b4.w));
      if (scale_res) {
        x.x = __fadd_rn(__fmul_rn(x.x, g4.x), rr[i].x); x.y = __fadd_rn(__fmul_rn(x.y, g4.y), rr[i].y);
        x.z = __fadd_rn(__fmul_rn(x.z, g4.z), rr[i].z); x.w = __fadd_rn(__fmul_rn(x.w, g4.w), rr[i].w);
      } else if (gelu_bwd) {
        x.x *= g2_gelu_grad(rr[i].x); x.y *= g2_gelu_grad(rr[i].y); x.z *= g2_gelu_grad(rr[i].z); x.w *= g2_gelu_grad(rr[i].w);
      } else if (has_res) {
        x.x = __fadd_rn(rr[i].x, x.x); x.y = __fadd_rn(rr[i].y, x.y);
        x.z = __fadd_rn(rr[i].z, x.z); x.w = __fadd_rn(rr[i].w, x.w);
      } else if (p.epi == LVAE_EPI_BIAS_GELU) {
        x.x = gelu_erf(x.x); x.y = gelu_erf(x.y); x.z = gelu_erf(x.z); x.w = gelu_erf(x.w);
      }
      const int64_t o = o0 + (int64_t)(8 * i) * p.N;
      if (p.out != nullptr) *reinterpret_cast<float4*>(p.out + o) = x;
      if (p.out_pl[0] != nullptr) {
        if (p.pl_act) { x.x = gelu_erf(x.x); x.y = gelu_erf(x.y); x.z = gelu_erf(x.z); x.w = gelu_erf(x.w); }
        float2 lo = make_float2(x.x, x.y), hi = make_float2(x.z, x.w);
        uint2 w;
        w.x = split_next(lo, f16); w.y = split_next(hi, f16);
        *reinterpret_cast<uint2*>(p.out_pl[0] + o) = w;
        if (p.out_pl[1] != nullptr) {
          w.x = split_next(lo, f16); w.y = split_next(hi, f16);
          *reinterpret_cast<uint2*>(p.out_pl[1] + o) = w;
        }
      }
    }
    __syncwarp();
  }
}

template <int EK>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2_THREADS, 1)
gemm2_tc_kernel(const __grid_constant__ G2Maps maps, const G2Params p) {
  extern __shared__ __align__(1024) uint8_t g2_smem_raw[];
  // the same padding in both CTAs of a pair (same kernel, same static layout): shared-memory descriptors are CTA-relative
  uint8_t* smem = g2_smem_raw + ((1024u - (smem_u32(g2_smem_raw) & 1023u)) & 1023u);
  const int a_tile = G2_BM * G2_BK * 2;                       // 16 KB: one plane of this CTA's 128 A rows
  const int b_tile = (p.BN >> 1) * G2_BK * 2;                 // one plane of this CTA's HALF of the W tile
  const int stage_bytes = 2 * (a_tile + b_tile);              // [A p0 | A p1 | W p0 | W p1]
  uint8_t* epi_smem = smem + (size_t)p.stages * stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + G2_EPI_WARPS * G2_STG);
  uint64_t* full_bar = bars;                                  // [stages]  used in the leader only (both CTAs' TMA bytes)
  uint64_t* empty_bar = bars + p.stages;                      // [stages]  per CTA, arrived by the leader's commit multicast
  uint64_t* tfull_bar = bars + 2 * p.stages;                  // [2]       per CTA, likewise
  uint64_t* tempty_bar = bars + 2 * p.stages + 2;             // [2]       leader only: 2 x 16 epilogue warps arrive
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * p.stages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
  const int nkb = (p.K + G2_BK - 1) / G2_BK;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(smem_u32(full_bar + s), 1); mbar_init(smem_u32(empty_bar + s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(smem_u32(tfull_bar + a), 1); mbar_init(smem_u32(tempty_bar + a), 2 * G2_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a[0]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.b[0]) : "memory");
  }
  if (warp == 1) {                                            // both CTAs, same warp: the pair's allocation
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();                                         // barriers + TMEM of BOTH CTAs exist before anyone signals
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ============================ TMA producer (both CTAs) ============================
    if (lane == 0) {
      const uint32_t full_leader = mapa_u32(smem_u32(full_bar), 0);
      int s = 0; uint32_t ph = 0;
      long long t_wait = 0; const long long t_begin = clock64();
      for (int pt = cid; pt < p.num_pair_tiles; pt += ncl) {
        const int mp = pt / p.n_tiles, nt = pt - mp * p.n_tiles;
        const int m0 = mp * (2 * G2_BM) + (int)rank * G2_BM;
        const int wrow0 = nt * p.BN + (int)rank * (p.BN >> 1);
        for (int kb = 0; kb < nkb; ++kb) {
          const long long tw = clock64();
          mbar_wait(smem_u32(empty_bar + s), ph ^ 1);
          t_wait += clock64() - tw;
          if (rank == 0) mbar_expect_tx(smem_u32(full_bar + s), (uint32_t)(2 * stage_bytes));   // both CTAs' bytes
          const uint32_t fb = full_leader + (uint32_t)(8 * s);
          const uint32_t base = smem_u32(smem + (size_t)s * stage_bytes);
          const int k0 = kb * G2_BK;
          const bool seg1 = k0 >= p.k_split;
#pragma unroll
          for (int pl = 0; pl < 2; ++pl) {
            tma_load_2d_cg2(base + pl * a_tile, seg1 ? &maps.a1[pl] : &maps.a[pl], fb, seg1 ? k0 - p.k_split : k0, m0);
            tma_load_2d_cg2(base + 2 * a_tile + pl * b_tile, &maps.b[pl], fb, k0, wrow0);
          }
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
      if (blockIdx.x == 0) { g2_prof[3] = clock64() - t_begin; g2_prof[4] = t_wait; }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer (leader CTA) ============================
    if (rank == 0) {
      const uint32_t fmt = p.f16 ? 0u : 1u;
      // D fp32, A / B 16-bit K-major, N = BN, M = 256 over the pair
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)((2 * G2_BM) >> 4) << 24);
      int s = 0; uint32_t ph = 0; int it = 0;
      long long t_full = 0, t_tempty = 0; const long long t_begin = clock64();
      for (int pt = cid; pt < p.num_pair_tiles; pt += ncl, ++it) {
        const int acc = p.nbuf == 2 ? (it & 1) : 0;
        const uint32_t use = p.nbuf == 2 ? ((uint32_t)it >> 1) : (uint32_t)it;
        { const long long tw = clock64(); mbar_wait(smem_u32(tempty_bar + acc), (use & 1) ^ 1); t_tempty += clock64() - tw; }
        tc_fence_after();
        const uint32_t d_main = tmem_base + (uint32_t)(acc * 2 * p.BN);
        const uint32_t d_cross = d_main + (uint32_t)p.BN;
        for (int kb = 0; kb < nkb; ++kb) {
          { const long long tw = clock64(); mbar_wait(smem_u32(full_bar + s), ph); t_full += clock64() - tw; }
          tc_fence_after();
          if (lane == 0) {
            const uint32_t base = smem_u32(smem + (size_t)s * stage_bytes);
            const uint64_t da0 = make_desc(base, G2_BK), da1 = make_desc(base + a_tile, G2_BK);
            const uint64_t db0 = make_desc(base + 2 * a_tile, G2_BK), db1 = make_desc(base + 2 * a_tile + b_tile, G2_BK);
#pragma unroll
            for (int k = 0; k < G2_BK / 16; ++k) {
              const uint64_t ko = (uint64_t)(k * 2);           // 16 elements = 32 bytes = 2 x 16-byte units along K
              const uint32_t first = (kb | k) ? 1u : 0u;
              tc_mma2(d_main, da0 + ko, db0 + ko, idesc, first);     // main  = a0 b0
              tc_mma2(d_cross, da0 + ko, db1 + ko, idesc, first);    // cross = a0 b1 + a1 b0 (same order as gemm_tc_kernel)
              tc_mma2(d_cross, da1 + ko, db0 + ko, idesc, 1u);
            }
            tc_commit2(smem_u32(empty_bar + s));               // frees the stage in both CTAs
            if (kb == nkb - 1) tc_commit2(smem_u32(tfull_bar + acc));
          }
          __syncwarp();
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
      if (blockIdx.x == 0 && lane == 0) { g2_prof[0] = clock64() - t_begin; g2_prof[1] = t_full; g2_prof[2] = t_tempty; g2_prof[8] = (unsigned long long)it; }
    }
  } else {
    // ============================ epilogue (both CTAs: own 128 rows x BN columns) ============================
    const int ew = warp - 2;
    const int q = warp & 3;                                    // TMEM lane quarter this warp may access
    const int cpar = ew >> 2;                                  // this warp takes the 32-column chunks cpar and cpar + 4
    uint32_t* stg = reinterpret_cast<uint32_t*>(epi_smem) + ew * (G2_STG / 4);
    const uint32_t tempty_leader = mapa_u32(smem_u32(tempty_bar), 0);
    const int nchunks = p.BN >> 5;
    int it = 0;
    long long t_tfull = 0, t_drain = 0; const long long t_begin = clock64();
    for (int pt = cid; pt < p.num_pair_tiles; pt += ncl, ++it) {
      const int acc = p.nbuf == 2 ? (it & 1) : 0;
      const uint32_t use = p.nbuf == 2 ? ((uint32_t)it >> 1) : (uint32_t)it;
      const int mp = pt / p.n_tiles, nt = pt - mp * p.n_tiles;
      const int row0 = mp * (2 * G2_BM) + (int)rank * G2_BM + q * 32;
      const int n0 = nt * p.BN;
      const long long tw = clock64();
      mbar_wait(smem_u32(tfull_bar + acc), use & 1);
      const long long td = clock64();
      t_tfull += td - tw;
      tc_fence_after();
      // drain: this warp's (up to) two chunks of main + cross into registers, then hand the accumulator back
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 2 * p.BN);
      const int c0 = cpar, c1 = cpar + 4;
      uint32_t s0[32], s1[32];
      g2_load_sum(tbase + (uint32_t)(c0 * 32), (uint32_t)p.BN, p.acc_scale, s0);          // BN >= 128: c0 < nchunks
      if (c1 < nchunks) g2_load_sum(tbase + (uint32_t)(c1 * 32), (uint32_t)p.BN, p.acc_scale, s1);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_leader + (uint32_t)(8 * acc));
      t_drain += clock64() - td;
      if (row0 < p.M) {
        if constexpr (EK == G2_GELU) {
          g2_chunk_gelu(p, s0, stg, row0, n0 + c0 * 32, lane);
          if (c1 < nchunks) g2_chunk_gelu(p, s1, stg, row0, n0 + c1 * 32, lane);
        } else {
          g2_chunk_rows(p, s0, stg, row0, n0 + c0 * 32, lane);
          if (c1 < nchunks) g2_chunk_rows(p, s1, stg, row0, n0 + c1 * 32, lane);
        }
      }
    }
    if (blockIdx.x == 0 && warp == 2 && lane == 0) { g2_prof[5] = clock64() - t_begin; g2_prof[6] = t_tfull; g2_prof[7] = t_drain; }
  }

  // nobody leaves while the peer may still read this CTA's shared memory / signal its barriers / use its TMEM
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------- host side
// Largest BN <= 256 with N % BN == 0 and BN % 64 == 0, BN >= 128 (0: this N does not tile); LVAE_G2_BN overrides.
static int g2_pick_bn(int N) {
  if (const char* e = getenv("LVAE_G2_BN")) {
    const int v = atoi(e);
    if (v >= 128 && v <= 256 && v % 32 == 0 && N % v == 0) return v;
  }
  for (int nt = (N + 255) / 256; nt <= N / 128; ++nt)
    if (N % nt == 0 && (N / nt) % 64 == 0 && N / nt <= 256) return N / nt;   // BN = 224 (7 x 32) is NOT bit-identical: the pair MMA sums N = 112 per CTA differently
  return 0;
}

static long long g2_launches = 0;      // diagnostics: how many GEMMs the pair kernel has taken (lvae_gemm2_launch_count)

// ek: 0 GELU (planes out), 1 ROWS.  *handled = 1 when the pair kernel took the GEMM, 0 when the caller should use
// gemm_tc_kernel (shape / size not eligible).  a_pl / a1_pl: [M, Ka] / [M, C1] planes; Ka = K, or C0 with a second segment.
int gemm2_tc_launch(const lvae_gemm_desc* d, const void* const* a_pl, const void* const* a1_pl, int M, int K, int Ka, int C1,
                    int ek, cudaStream_t stream, int* handled) {
  *handled = 0;
  // OPT-IN (LVAE_GEMM2=1): measured on B200 (profiles/r2_gemm2_pair_kernel.md) the pair kernel is 10-25 % SLOWER than
  // gemm_tc_kernel on every qarv shape -- BN = 256 / 192 run the MMAs at their nominal rate but expose the 16 BN-cycle
  // TMEM drain (64 B/clk) of the single accumulator buffer; BN = 128 double-buffers but then moves as many shared-memory
  // bytes per MMA cycle as the 128-row kernel (the N-doubled a0 [b0 | b1] instruction does not combine with the pair's
  // split B operand).  Kept as the measured baseline for that design, off by default.
  { const char* e = getenv("LVAE_GEMM2"); if (!e || atoi(e) == 0) return 0; }
  const bool two_planes = d->precision == LVAE_PREC_F16X3 || d->precision == LVAE_PREC_BF16X3;
  if (!two_planes || K < 64 || K % 8 != 0 || d->N % 4 != 0) return 0;
  const int BN = g2_pick_bn(d->N);
  if (BN == 0) return 0;
  const int n_tiles = d->N / BN;
  const int pair_tiles = ((M + 2 * G2_BM - 1) / (2 * G2_BM)) * n_tiles;
  int min_tiles = 64;                                          // below ~one wave of pairs the 128-row kernel fills the chip better
  { const char* e = getenv("LVAE_G2_MIN_TILES"); if (e) min_tiles = atoi(e); }
  if (pair_tiles < min_tiles) return 0;
  if (a1_pl != nullptr && (Ka % G2_BK != 0 || C1 % 8 != 0)) return 0;

  G2Params p;
  p.M = M; p.N = d->N; p.K = K; p.BN = BN; p.n_tiles = n_tiles; p.num_pair_tiles = pair_tiles;
  p.nbuf = 4 * BN <= 512 ? 2 : 1;
  p.k_split = a1_pl ? Ka : K;
  p.bias = d->bias; p.gamma = d->gamma; p.res = d->res; p.out = d->out;
  p.out_pl[0] = (uint16_t*)d->out_planes[0];
  p.out_pl[1] = p.out_pl[0] ? (uint16_t*)d->out_planes[1] : nullptr;
  p.epi = d->epilogue; p.pl_act = d->out_planes_act;
  p.f16 = d->precision == LVAE_PREC_F16X3 ? 1 : 0;
  p.acc_scale = p.f16 ? 1.0f / LVAE_F16_WEIGHT_SCALE : 1.0f;
  const int stage_bytes = 2 * (G2_BM * G2_BK * 2 + (BN / 2) * G2_BK * 2);
  const int fixed = 1024 + G2_EPI_WARPS * G2_STG + 256;
  int stages = (227 * 1024 - fixed) / stage_bytes;
  const int nkb = (K + G2_BK - 1) / G2_BK;
  if (stages > 8) stages = 8;
  if (stages > nkb + 1) stages = nkb + 1;
  if (stages < 2) return 0;
  p.stages = stages;

  G2Maps maps;
  int rc;
  for (int i = 0; i < 2; ++i) {
    if ((rc = make_map(&maps.a[i], a_pl[i], M, Ka, G2_BM, G2_BK))) return rc;
    if ((rc = make_map(&maps.b[i], d->w_planes[i], d->N, K, BN / 2, G2_BK))) return rc;
    if (a1_pl) { if ((rc = make_map(&maps.a1[i], a1_pl[i], M, C1, G2_BM, G2_BK))) return rc; }
    else maps.a1[i] = maps.a[i];
  }
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    LVAE_CUDA_CALL(cudaFuncSetAttribute(gemm2_tc_kernel<G2_GELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    LVAE_CUDA_CALL(cudaFuncSetAttribute(gemm2_tc_kernel<G2_ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  const int smem = fixed + stages * stage_bytes;
  const int ncl = pair_tiles < n_sm / 2 ? pair_tiles : n_sm / 2;
  if (ek == G2_GELU) gemm2_tc_kernel<G2_GELU><<<2 * ncl, G2_THREADS, smem, stream>>>(maps, p);
  else gemm2_tc_kernel<G2_ROWS><<<2 * ncl, G2_THREADS, smem, stream>>>(maps, p);
  LVAE_CUDA_LAUNCH_CHECK();
  *handled = 1;
  ++g2_launches;
  return 0;
}

}  // namespace lvae

extern "C" long long lvae_gemm2_launch_count(void) { return lvae::g2_launches; }

namespace lvae { int gemm_tc_prof_read(unsigned long long* out16); int gemm_tc_set_tuning(int which, int value); }
extern "C" int lvae_set_tuning(int which, int value) { return lvae::gemm_tc_set_tuning(which, value); }
extern "C" int lvae_debug_prof(int which, unsigned long long* out16) {
  if (which == 2) {
    LVAE_CUDA_CALL(cudaDeviceSynchronize());
    LVAE_CUDA_CALL(cudaMemcpyFromSymbol(out16, lvae::g2_prof, sizeof(unsigned long long) * 16));
    return 0;
  }
  return lvae::gemm_tc_prof_read(out16);
}
