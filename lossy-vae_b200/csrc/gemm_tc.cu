// tcgen05 tensor-core GEMM for the dense contractions of the path (LVAE_PREC_BF16X3 / LVAE_PREC_BF16).
//
//   out[m, n] = epilogue( sum_k A[m, k] * W[n, k] + bias[n] ),   A: [M, K], W: [N, K], both K-contiguous
//
// Operands are bf16 PLANES: an fp32 value x is carried as p0 = rn(x), p1 = rn(x - p0), p2 = rn(x - p0 - p1)
// (24 mantissa bits in three 8-bit pieces).  Per logical product, with fp32 accumulation in TMEM:
//   BF16    1 plane,  1 MMA   p0*p0                                        ~2^-8
//   BF16X3  2 planes, 3 products p0*p0 + p0*p1 + p1*p0 (2 MMAs: p0 * [p0; p1] is one)   ~2^-17
//   BF16X6  3 planes, 6 MMAs  ... + p1*p1 + p0*p2 + p2*p0                  ~2^-23 (fp32-class: the mode in which
//                                                                          the quantised symbols match the CPU oracle)
//
// Kernel structure (one persistent CTA per SM, 6 warps, no clusters):
//   warp 0   TMA producer: cp.async.bulk.tensor.2d of the [128 x 64] A tile(s) and [BN x 64] W tile(s) of a
//            k-block into a SWIZZLE_128B ring of shared-memory stages, completion on an mbarrier
//   warp 1   allocates TMEM, then one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M = 128, N = BN,
//            K = 16 per instruction) from shared-memory descriptors; tcgen05.commit releases the stage and,
//            after the last k-block, publishes the accumulator
//   warps 2-5  epilogue: tcgen05.ld the 128 x BN fp32 accumulator (32 lanes per warp), transpose 32 x 32
//            chunks through padded shared memory so that global traffic is 128-byte coalesced, apply
//            bias / GELU / layer-scale + residual / pixel-shuffle, write fp32 and/or bf16 hi-lo planes
// Accumulation: tcgen05.mma adds into the fp32 TMEM accumulator with truncation, one rounding per instruction,
// so the error of an accumulator grows with the number of MMAs chained into it (measured: bf16x6 with all six
// terms in one accumulator was no more accurate than bf16x3).  The multi-plane modes therefore keep TWO
// accumulators per tile: `main` receives only the p0*p0 products (K/16 roundings), `cross` all correction
// terms, whose magnitude -- and truncation error -- is 2^-8 of the main sum; the epilogue adds them once (RN).
// Two such buffers in TMEM (2 x 2 x BN <= 512 columns, BN <= 128; 2 x BN with BN <= 256 for single-plane bf16)
// let the epilogue of tile i overlap the MMAs of tile i+1.  K order is fixed, there is no split-K and BN depends on N only, so results are bit-reproducible
// and do not depend on the batch an image travels in.
#include "tc_common.cuh"
#include "latent_math.cuh"
#include <cuda_bf16.h>
#include <mutex>
#include <stdlib.h>

namespace lvae {

constexpr int TC_BM = 128;
// Epilogue specialisations (template parameter EK): each instantiation carries only its own epilogue code -- the
// all-in-one kernel was 9 k SASS instructions and spent 27-32 % of its stall samples on instruction fetch.
//   EK_GELU  fc1: bias + GELU -> 16-bit planes.  CUDA-core bound (17 instructions per element for the ATen erf
//            formula alone), so it runs 16 epilogue warps: four per TMEM lane quarter, one 32 x 32 chunk each
//   EK_ROWS  bias | layer-scale + residual | bias + residual -> fp32 rows (+ planes), 128-bit accesses
//   EK_MISC  pixel-shuffle stores, implicit-conv pixel tiles, odd N: scalar lane = column path
constexpr int EK_GELU = 0, EK_ROWS = 1, EK_MISC = 2;
#ifndef LVAE_TC_GELU_WARPS
#define LVAE_TC_GELU_WARPS 16
#endif
__host__ __device__ constexpr int tc_epi_warps(int ek) { return ek == EK_GELU ? LVAE_TC_GELU_WARPS : 8; }
__host__ __device__ constexpr int tc_threads(int ek) { return 64 + 32 * tc_epi_warps(ek); }
// per-warp transpose buffer in words: GELU 32 rows x 64 B (one 16-bit plane of 32 columns), ROWS 32 rows x 128 B --
// both XOR-swizzled at 16-byte granularity instead of padded (conflict-free 128-bit row writes and row-segment
// reads; padding would cost the third 64 KB pipeline stage at BN = 128) -- MISC 32 x 33 words
__host__ __device__ constexpr int tc_stage_words(int ek) { return ek == EK_GELU ? 32 * 16 : (ek == EK_ROWS ? 32 * 32 : 32 * 33); }
__host__ __device__ constexpr int tc_epi_stage_bytes(int ek) { return tc_epi_warps(ek) * tc_stage_words(ek) * 4; }

struct TcParams {
  int M, N, K;
  int BN, BK, n_tiles, num_tiles, stages, tmem_cols, acc_cols;   // acc_cols: TMEM columns per tile buffer (BN or 2*BN)   // BK: bf16 elements per k-block = one swizzle row (64 -> 128 B, 32 -> 64 B)
  const float* bias; const float* gamma; const float* res;
  float* out; __nv_bfloat16* out_pl[3];
  int epi, r, Ho, Wo;
  int pl_act;            // out planes hold gelu(result) (EK_ROWS)
  int k_split;           // k >= k_split comes from the a1 maps (K-concat of two plane sets); = K when there is one segment
  int prefetch;          // L2-prefetch the next tile's A k-blocks (tuning knob LVAE_TC_PREFETCH=1; measured: no gain --
                         // the large-K GEMMs sit on the ~6.3 kB/clk L2->SM cap (85 B/clk/SM wanted at BN = 128), not on latency)
  int f16;               // plane element format: 0 bf16, 1 fp16 (LVAE_PREC_F16X3)
  float acc_scale;       // 1 / (scale the weight planes carry): 1, or 2^-8 in the fp16 mode -- exact
  // implicit 3x3 conv (stride 1, pad 1) from NHWC planes: an M-tile is a CONV_TH x CONV_TW pixel patch of one image
  int conv, cH, cW, cC, tiles_w, tiles_h;
  // latent epilogue of the implicit 3x3 posterior convolution (lvae_gemm_latent): the tile's result is qm; with the prior
  // parameters read from lat_prior [M, 2 N] the epilogue quantises, evaluates the likelihood and writes z (p.out), the per-image
  // rate partials (one slot per (pixel tile, 32-column chunk, row quarter): lat_kl[image * lat_kl_stride + slot]) and -- for the
  // coder -- symbols and table indexes in NCHW order.  nullptr: plain epilogue.
  const float* lat_prior; float* lat_kl; long long lat_kl_stride; float* lat_kl_elem;
  int32_t* lat_sym; int32_t* lat_idx; const float* lat_table; int lat_nscales, lat_cdf;
  // split-K (weight gradients: K = pixels is the long dimension): tile index t = z * base_tiles + tile, split z owns the
  // k-blocks [z * kb_per, (z + 1) * kb_per); partial results meet in fp32 atomics on `out` (zeroed by the caller)
  int base_tiles, kb_per, atomic;
};
constexpr int CONV_TW = 16, CONV_TH = 8;       // 128 output pixels per tile

struct TcMaps { CUtensorMap a[3]; CUtensorMap b[3]; CUtensorMap a1[3]; };   // a1: second K segment (K-concat from planes)

// cycle breakdown of CTA 0 (diagnostics, lvae_debug_prof(1, ...)): [0] MMA thread total, [1] waiting for operands (full),
// [2] waiting for the accumulator (tempty), [3] producer total, [4] producer waiting for a free stage, [8] tiles
__device__ unsigned long long tc_prof[16];
int gemm_tc_prof_read(unsigned long long* out16) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpyFromSymbol(out16, tc_prof, sizeof(unsigned long long) * 16);
  return (int)e;
}

// d/dh gelu(h) = Phi(h) + h phi(h)  (erf form, as F.gelu's autograd)
__device__ __forceinline__ float gelu_grad(float h) {
  const float cdf = 0.5f * (1.f + erff(h * 0.70710678118654752f));
  return fmaf(h * 0.39894228040143268f, __expf(-0.5f * h * h), cdf);
}

__device__ __forceinline__ float epi_value(const TcParams& p, int m, int n, float acc) {
  float v = acc;
  if (p.bias) v = __fadd_rn(v, __ldg(p.bias + n));
  switch (p.epi) {
    case LVAE_EPI_BIAS_GELU: v = gelu_erf(v); break;
    case LVAE_EPI_SCALE_RES: v = __fadd_rn(__fmul_rn(v, __ldg(p.gamma + n)), p.res[(int64_t)m * p.N + n]); break;
    case LVAE_EPI_BIAS_RES:  v = __fadd_rn(p.res[(int64_t)m * p.N + n], v); break;
    case LVAE_EPI_GELU_BWD:  v *= gelu_grad(p.res[(int64_t)m * p.N + n]); break;
    default: break;
  }
  return v;
}

__device__ __forceinline__ void store_planes(const TcParams& p, int64_t o, float v) {
  float2 r = make_float2(v, 0.f);
  const bool f16 = p.f16 != 0;
  reinterpret_cast<uint16_t*>(p.out_pl[0])[o] = (uint16_t)split_next(r, f16);
  if (p.out_pl[1]) {
    reinterpret_cast<uint16_t*>(p.out_pl[1])[o] = (uint16_t)split_next(r, f16);
    if (p.out_pl[2]) reinterpret_cast<uint16_t*>(p.out_pl[2])[o] = (uint16_t)split_next(r, f16);
  }
}

__device__ __forceinline__ void epi_store(const TcParams& p, int m, int n, float v) {
  if (p.epi == LVAE_EPI_SHUFFLE_NHWC || p.epi == LVAE_EPI_SHUFFLE_NCHW) {
    const int r = p.r, Co = p.N / (r * r);
    const int q = n / Co, c = n - q * Co, i = q / r, j = q - i * r;
    const int wo = m % p.Wo; const int t = m / p.Wo; const int ho = t % p.Ho; const int b = t / p.Ho;
    const int Hr = p.Ho * r, Wr = p.Wo * r;
    if (p.epi == LVAE_EPI_SHUFFLE_NHWC) p.out[(((int64_t)b * Hr + ho * r + i) * Wr + wo * r + j) * Co + c] = v;
    else p.out[(((int64_t)b * Co + c) * Hr + ho * r + i) * Wr + wo * r + j] = v;
    return;
  }
  const int64_t o = (int64_t)m * p.N + n;
  if (p.out) p.out[o] = v;
  if (p.out_pl[0]) store_planes(p, o, v);
}

template <int NPL, int EK>
__global__ void __launch_bounds__(tc_threads(EK), 1)
gemm_tc_kernel(const __grid_constant__ TcMaps maps, const TcParams p) {
  constexpr int TC_EPI_WARPS = tc_epi_warps(EK);
  constexpr int TC_EPI_STAGE = tc_epi_stage_bytes(EK);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment of the swizzled tiles
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // (pointer arithmetic keeps the shared address space)
  const int a_tile = TC_BM * p.BK * 2;
  const int b_tile = p.BN * p.BK * 2;
  const int stage_bytes = NPL * (a_tile + b_tile);          // [A planes | B planes]
  uint8_t* epi_smem = smem + (size_t)p.stages * stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + TC_EPI_STAGE);
  uint64_t* full_bar = bars;                                // [stages]
  uint64_t* empty_bar = bars + p.stages;                    // [stages]
  uint64_t* tfull_bar = bars + 2 * p.stages;                // [2]
  uint64_t* tempty_bar = bars + 2 * p.stages + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * p.stages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (p.K + p.BK - 1) / p.BK;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(smem_u32(full_bar + s), 1); mbar_init(smem_u32(empty_bar + s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(smem_u32(tfull_bar + a), 1); mbar_init(smem_u32(tempty_bar + a), TC_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a[0]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.b[0]) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // PDL: everything above overlapped the predecessor's tail; from here on this kernel reads what it wrote
  pdl_trigger();
  pdl_wait();

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      long long t_wait = 0; const long long t_begin = clock64();
      for (int tz = blockIdx.x; tz < p.num_tiles; tz += gridDim.x) {
        const int z = tz / p.base_tiles, t = tz - z * p.base_tiles;
        const int kb0 = z * p.kb_per, kb1 = min(nkb, kb0 + p.kb_per);
        const int mt = t / p.n_tiles;
        const int m0 = mt * TC_BM, n0 = (t % p.n_tiles) * p.BN;
        // conv mode: tile -> (image, patch row, patch column)
        const int ctx = mt % p.tiles_w, cty = (mt / p.tiles_w) % p.tiles_h, cb = mt / (p.tiles_w * p.tiles_h);
        const int cpt = p.cC / p.BK;                             // k-blocks per filter tap
        for (int kb = kb0; kb < kb1; ++kb) {
          { const long long tw = clock64(); mbar_wait(smem_u32(empty_bar + s), ph ^ 1); t_wait += clock64() - tw; }
          const uint32_t fb = smem_u32(full_bar + s);
          mbar_expect_tx(fb, (uint32_t)stage_bytes);
          const uint32_t base = smem_u32(smem + (size_t)s * stage_bytes);
          if (p.conv) {
            // the (dy, dx)-shifted patch of the input planes: out-of-image pixels are zero-filled by TMA (= padding)
            const int tap = kb / cpt, c0 = (kb - tap * cpt) * p.BK;
            const int dy = tap / 3 - 1, dx = tap % 3 - 1;
#pragma unroll
            for (int pl = 0; pl < NPL; ++pl) {
              tma_load_4d(base + pl * a_tile, &maps.a[pl], fb, c0, ctx * CONV_TW + dx, cty * CONV_TH + dy, cb);
              tma_load_2d(base + NPL * a_tile + pl * b_tile, &maps.b[pl], fb, kb * p.BK, n0);
            }
          } else {
#pragma unroll
            const int k0 = kb * p.BK;
            const bool seg1 = k0 >= p.k_split;
            for (int pl = 0; pl < NPL; ++pl) {
              tma_load_2d(base + pl * a_tile, seg1 ? &maps.a1[pl] : &maps.a[pl], fb, seg1 ? k0 - p.k_split : k0, m0);
              tma_load_2d(base + NPL * a_tile + pl * b_tile, &maps.b[pl], fb, k0, n0);
            }
            // optional: pull the same k-block of this CTA's NEXT tile into L2 (off by default, see TcParams::prefetch)
            if (p.prefetch) {
              const int tn = t + gridDim.x;
              if (tn < p.base_tiles && (tn / p.n_tiles) != mt) {
                const int m1 = (tn / p.n_tiles) * TC_BM;
#pragma unroll
                for (int pl = 0; pl < NPL; ++pl) tma_prefetch_2d(&maps.a[pl], kb * p.BK, m1);
              }
            }
          }
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
      if (blockIdx.x == 0) { tc_prof[3] = clock64() - t_begin; tc_prof[4] = t_wait; }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    // instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): D fp32, A/B bf16, both K-major
    // a_format / b_format (bits 7-9 / 10-12): 0 = fp16, 1 = bf16
    const uint32_t fmt = p.f16 ? 0u : 1u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    const uint32_t idesc2n = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.BN >> 2) << 17) | ((uint32_t)(TC_BM >> 4) << 24);   // N = 2 * BN
    int s = 0; uint32_t ph = 0; int it = 0;
    long long t_full = 0, t_tempty = 0; const long long t_begin = clock64();
    for (int tz = blockIdx.x; tz < p.num_tiles; tz += gridDim.x, ++it) {
      const int acc = it & 1;
      const int kb0 = (tz / p.base_tiles) * p.kb_per, kb1 = min(nkb, kb0 + p.kb_per);
      { const long long tw = clock64(); mbar_wait(smem_u32(tempty_bar + acc), (((uint32_t)it >> 1) & 1) ^ 1); t_tempty += clock64() - tw; }
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.acc_cols);      // main accumulator
      const uint32_t d_cross = d_tmem + (uint32_t)p.BN;                       // correction terms (NPL >= 2)
      for (int kb = kb0; kb < kb1; ++kb) {
        { const long long tw = clock64(); mbar_wait(smem_u32(full_bar + s), ph); t_full += clock64() - tw; }
        tc_fence_after();
        if (lane == 0) {
          const uint32_t base = smem_u32(smem + (size_t)s * stage_bytes);
          uint64_t da[NPL], db[NPL];
#pragma unroll
          for (int pl = 0; pl < NPL; ++pl) {
            da[pl] = make_desc(base + pl * a_tile, p.BK);
            db[pl] = make_desc(base + NPL * a_tile + pl * b_tile, p.BK);
          }
          const int ksteps = p.BK / 16;
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t ko = (uint64_t)(k * 2);           // 16 bf16 = 32 bytes = 2 x 16-byte units along K
            const uint32_t first = ((kb - kb0) | k) ? 1u : 0u;
            if (NPL >= 2) {
              // one tcgen05.mma costs ~142 cycles whatever its N (scripts/probe/mma_probe.cu), so a0 * b0 (-> main) and
              // a0 * b1 (-> cross) ride in ONE instruction of N = 2 * BN: the B planes 0 and 1 are adjacent in the stage
              // and main | cross are adjacent in TMEM.  2 MMAs per k-step instead of 3 (5 instead of 6 with 3 planes).
              tc_mma(d_tmem, da[0] + ko, db[0] + ko, idesc2n, first);
              tc_mma(d_cross, da[1] + ko, db[0] + ko, idesc, 1u);
              if (NPL == 3) {
                tc_mma(d_cross, da[1] + ko, db[1] + ko, idesc, 1u);
                tc_mma(d_cross, da[2] + ko, db[0] + ko, idesc, 1u);
                tc_mma(d_cross, da[0] + ko, db[2] + ko, idesc, 1u);
              }
            } else {
              tc_mma(d_tmem, da[0] + ko, db[0] + ko, idesc, first);
            }
          }
          tc_commit(smem_u32(empty_bar + s));               // frees the stage once these MMAs have read it
          if (kb == kb1 - 1) tc_commit(smem_u32(tfull_bar + acc));
        }
        __syncwarp();
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
    }
    if (blockIdx.x == 0 && lane == 0) { tc_prof[0] = clock64() - t_begin; tc_prof[1] = t_full; tc_prof[2] = t_tempty; tc_prof[8] = (unsigned long long)it; }
  } else {
    // ============================ epilogue (warps 2 .. 2+TC_EPI_WARPS) ============================
    const int ew = warp - 2;
    const int q = warp & 3;                                  // TMEM lane quarter this warp may access
    const int cpar = ew >> 2, cstep = TC_EPI_WARPS / 4;      // this warp takes chunks cpar, cpar + cstep, ...
    uint32_t* stg = reinterpret_cast<uint32_t*>(epi_smem) + ew * tc_stage_words(EK);
    const bool f16 = p.f16 != 0;
    int it = 0;
    for (int tz = blockIdx.x; tz < p.num_tiles; tz += gridDim.x, ++it) {
      const int acc = it & 1;
      const int t = tz % p.base_tiles;
      const int m0 = (t / p.n_tiles) * TC_BM, n0 = (t % p.n_tiles) * p.BN;
      mbar_wait(smem_u32(tfull_bar + acc), ((uint32_t)it >> 1) & 1);
      tc_fence_after();
      const int row0 = m0 + q * 32;
      const int nchunks = (p.BN + 31) / 32;
      int last_c = cpar; while (last_c + cstep < nchunks) last_c += cstep;
      if (cpar >= nchunks) {                                 // narrow tile: nothing for this warp to read
        if (lane == 0) mbar_arrive(smem_u32(tempty_bar + acc));
        continue;
      }
      for (int c = cpar; c < nchunks; c += cstep) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.acc_cols + c * 32);
        const int width = (p.BN - c * 32) >= 32 ? 32 : 16;   // BN is a multiple of 16
        if (NPL >= 2) {
          uint32_t u[32];
          if (width == 32) { tc_ld32(taddr, v); tc_ld32(taddr + (uint32_t)p.BN, u); }
          else { tc_ld16(taddr, v); tc_ld16(taddr + (uint32_t)p.BN, u); }
          tc_wait_ld();
          const float2 sc = splat2(p.acc_scale);
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            // (main + cross) * 2^-s: the scale is a power of two, so the product is exact
            const float2 sum = mul2(add2(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])),
                                         make_float2(__uint_as_float(u[j]), __uint_as_float(u[j + 1]))), sc);
            const bool ok = (j < 16 || width == 32);
            v[j] = ok ? __float_as_uint(sum.x) : 0u;
            v[j + 1] = ok ? __float_as_uint(sum.y) : 0u;
          }
        } else {
          if (width == 32) tc_ld32(taddr, v);
          else {
            tc_ld16(taddr, v);
#pragma unroll
            for (int j = 16; j < 32; ++j) v[j] = 0u;
          }
          tc_wait_ld();
          if (p.acc_scale != 1.0f) {                         // single fp16 plane (LVAE_PREC_F16): weights carry 2^8
            const float2 sc = splat2(p.acc_scale);
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float2 x = mul2(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), sc);
              v[j] = __float_as_uint(x.x); v[j + 1] = __float_as_uint(x.y);
            }
          }
        }
        if (c == last_c) {                                   // this warp has read its share of the accumulator
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(tempty_bar + acc));
        }
        const int nb = n0 + c * 32;                          // first column of the chunk
        if constexpr (EK == EK_GELU) {
          // ---- fc1: bias + GELU in the row-owner layout (16 independent pair chains per thread), then per plane
          //      pack 16-bit pairs, transpose through shared memory (80-byte rows: conflict-free 128-bit stores),
          //      and store 64-byte row segments with 64-bit accesses
          if (nb + 32 <= p.N && p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float2 b2 = __ldg(reinterpret_cast<const float2*>(p.bias + nb + j));
              const float2 o = gelu_erf2(add2(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), b2));
              v[j] = __float_as_uint(o.x); v[j + 1] = __float_as_uint(o.y);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int n = nb + j;
              const float b = (p.bias != nullptr && n < p.N) ? __ldg(p.bias + n) : 0.f;
              v[j] = __float_as_uint(gelu_erf(__fadd_rn(__uint_as_float(v[j]), b)));
            }
          }
          const int sub = lane >> 3, l8 = lane & 7;          // 4 rows per pass, 8 lanes x 4 columns per row
          const int n = nb + 4 * l8;
          const bool full = (row0 + 32 <= p.M) && (nb + 32 <= p.N) && (width == 32);
#pragma unroll
          for (int pl = 0; pl < NPL; ++pl) {
            if (p.out_pl[pl] == nullptr) break;
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              uint32_t w[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int j = j4 * 8 + e * 2;
                float2 gv = make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1]));
                if (pl + 1 < NPL) {                          // leaves the exact residual for the next plane
                  w[e] = split_next(gv, f16);
                  v[j] = __float_as_uint(gv.x); v[j + 1] = __float_as_uint(gv.y);
                } else {
                  w[e] = f16 ? pack2<true>(gv.x, gv.y) : pack2<false>(gv.x, gv.y);
                }
              }
              *reinterpret_cast<uint4*>(stg + lane * 16 + ((j4 ^ ((lane >> 1) & 3)) << 2)) = make_uint4(w[0], w[1], w[2], w[3]);
            }
            __syncwarp();
            __nv_bfloat16* dst = p.out_pl[pl] + (int64_t)(row0 + sub) * p.N + n;
            if (full) {
#pragma unroll
              for (int r = 0; r < 32; r += 4)
                *reinterpret_cast<uint2*>(dst + (int64_t)r * p.N) =
                    *reinterpret_cast<const uint2*>(stg + (r + sub) * 16 + ((((l8 >> 1) ^ (((r + sub) >> 1) & 3)) << 2) | ((l8 & 1) << 1)));
            } else {
              for (int r = 0; r < 32; r += 4) {
                const uint2 w2 = *reinterpret_cast<const uint2*>(stg + (r + sub) * 16 + ((((l8 >> 1) ^ (((r + sub) >> 1) & 3)) << 2) | ((l8 & 1) << 1)));
                if (row0 + r + sub < p.M && 4 * l8 < width) {
                  if (n + 3 < p.N) *reinterpret_cast<uint2*>(dst + (int64_t)r * p.N) = w2;
                  else if (n + 1 < p.N) *reinterpret_cast<uint32_t*>(dst + (int64_t)r * p.N) = w2.x;
                }
              }
            }
            __syncwarp();
          }
        } else if constexpr (EK == EK_ROWS) {
          // ---- bias | layer-scale + residual | bias + residual -> fp32 rows (+ planes).  The row-owner registers
          //      go through shared memory as 128-bit rows (XOR-swizzled: conflict-free both ways); then lane =
          //      4 consecutive columns, 4 rows per pass, all residual loads in flight before the first is consumed
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4)
            *reinterpret_cast<uint4*>(stg + lane * 32 + ((j4 ^ (lane & 7)) << 2)) = make_uint4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
          __syncwarp();
          const int sub = lane >> 3, l8 = lane & 7;
          const int n = nb + 4 * l8;
          if (4 * l8 < width && n < p.N) {                   // N % 4 == 0 (host-checked): n + 3 < N
            const float4 b4 = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
            const bool scale_res = p.epi == LVAE_EPI_SCALE_RES;
            const bool gelu_bwd = p.epi == LVAE_EPI_GELU_BWD;
            const bool has_res = scale_res || gelu_bwd || p.epi == LVAE_EPI_BIAS_RES;
            const float4 g4 = scale_res ? __ldg(reinterpret_cast<const float4*>(p.gamma + n)) : make_float4(1.f, 1.f, 1.f, 1.f);
            const int64_t o0 = (int64_t)(row0 + sub) * p.N + n;
            const int rows = p.M - row0 - sub;               // valid while 4 * i < rows
            float4 rr[8];
            if (has_res) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                rr[i] = (4 * i < rows) ? *reinterpret_cast<const float4*>(p.res + o0 + (int64_t)(4 * i) * p.N) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (4 * i >= rows) break;
              const float4 a = *reinterpret_cast<const float4*>(stg + (4 * i + sub) * 32 + ((l8 ^ ((4 * i + sub) & 7)) << 2));
              float4 x = make_float4(__fadd_rn(a.x, b4.x), __fadd_rn(a.y, b4.y), __fadd_rn(a.z, b4.z), __fadd_rn(a.w, b4.w));
              if (scale_res) {
                x.x = __fadd_rn(__fmul_rn(x.x, g4.x), rr[i].x); x.y = __fadd_rn(__fmul_rn(x.y, g4.y), rr[i].y);
                x.z = __fadd_rn(__fmul_rn(x.z, g4.z), rr[i].z); x.w = __fadd_rn(__fmul_rn(x.w, g4.w), rr[i].w);
              } else if (gelu_bwd) {
                x.x *= gelu_grad(rr[i].x); x.y *= gelu_grad(rr[i].y); x.z *= gelu_grad(rr[i].z); x.w *= gelu_grad(rr[i].w);
              } else if (has_res) {
                x.x = __fadd_rn(rr[i].x, x.x); x.y = __fadd_rn(rr[i].y, x.y);
                x.z = __fadd_rn(rr[i].z, x.z); x.w = __fadd_rn(rr[i].w, x.w);
              } else if (p.epi == LVAE_EPI_BIAS_GELU) {
                x.x = gelu_erf(x.x); x.y = gelu_erf(x.y); x.z = gelu_erf(x.z); x.w = gelu_erf(x.w);
              }
              const int64_t o = o0 + (int64_t)(4 * i) * p.N;
              if (p.out != nullptr) *reinterpret_cast<float4*>(p.out + o) = x;
              if (p.out_pl[0] != nullptr) {
                if (p.pl_act) { x.x = gelu_erf(x.x); x.y = gelu_erf(x.y); x.z = gelu_erf(x.z); x.w = gelu_erf(x.w); }
                float2 lo = make_float2(x.x, x.y), hi = make_float2(x.z, x.w);
#pragma unroll
                for (int pl = 0; pl < NPL; ++pl) {
                  if (p.out_pl[pl] == nullptr) break;
                  uint2 w;
                  w.x = split_next(lo, f16); w.y = split_next(hi, f16);
                  *reinterpret_cast<uint2*>(p.out_pl[pl] + o) = w;
                }
              }
            }
          }
          __syncwarp();
        } else {
          // ---- shuffle stores, implicit-conv pixel tiles, odd N: transpose, then lane = column
#pragma unroll
          for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = v[j];
          __syncwarp();
          const int n = nb + lane;
          const bool n_ok = (lane < width) && (n < p.N);
          if (p.conv && p.lat_prior != nullptr) {
            // ---- latent epilogue: qm -> (z, -ln P, symbol, table index), arithmetic of latent_math.cuh (= latent_kernel<eval>)
            const int mt = t / p.n_tiles;
            const int ctx = mt % p.tiles_w, cty = (mt / p.tiles_w) % p.tiles_h, cb = mt / (p.tiles_w * p.tiles_h);
            const int hw_img = p.cH * p.cW;
            // the chunk holds 32 pixels x wv valid columns (wv = zdim = 8 on the largest layers): element i = lane + 32 k is
            // pixel i / wv, column i % wv, so that every lane works whatever the width (lane = column would leave 24 of 32 idle)
            int wv = p.N - nb; wv = wv > width ? width : wv; wv = wv < 0 ? 0 : wv;
            float kl_sum = 0.f;
            for (int i = lane; i < 32 * wv; i += 32) {
              const int r = i / wv, cn = i - r * wv;
              const int row = q * 32 + r;
              const int hh = cty * CONV_TH + row / CONV_TW, ww = ctx * CONV_TW + row % CONV_TW;
              if (hh < p.cH && ww < p.cW) {
                const int nn = nb + cn;
                const float qv = __fadd_rn(__uint_as_float(stg[r * 33 + cn]), p.bias ? __ldg(p.bias + nn) : 0.f);
                const int64_t m = ((int64_t)cb * p.cH + hh) * p.cW + ww;
                const float pm = __ldg(p.lat_prior + m * 2 * p.N + nn), pl = __ldg(p.lat_prior + m * 2 * p.N + p.N + nn);
                float zz, kl, rr, sc;
                latent_elem<0>(qv, pm, pl, 0.f, p.lat_cdf, zz, kl, rr, sc);
                p.out[m * p.N + nn] = zz;
                if (p.lat_kl_elem != nullptr) p.lat_kl_elem[m * p.N + nn] = kl;
                if (p.lat_sym != nullptr) {
                  const int64_t o = ((int64_t)cb * p.N + nn) * hw_img + (int64_t)hh * p.cW + ww;     // NCHW order for the coder
                  p.lat_sym[o] = (int32_t)rr;
                  p.lat_idx[o] = scale_index(sc, p.lat_table, p.lat_nscales);
                }
                kl_sum += kl;
              }
            }
            kl_sum = warp_sum(kl_sum);
            if (lane == 0) {
              const int nch = (p.N + 31) / 32;
              const int slot = (((cty * p.tiles_w + ctx) * nch + (nb >> 5)) << 2) + q;
              p.lat_kl[(int64_t)cb * p.lat_kl_stride + slot] = kl_sum;
            }
          } else if (n_ok && p.conv) {
            // rows of the tile are the pixels of a CONV_TH x CONV_TW patch: row -> (h, w) -> m; N is small here
            const int mt = t / p.n_tiles;
            const int ctx = mt % p.tiles_w, cty = (mt / p.tiles_w) % p.tiles_h, cb = mt / (p.tiles_w * p.tiles_h);
            const float b = p.bias ? __ldg(p.bias + n) : 0.f;
            const bool gelu = p.epi == LVAE_EPI_BIAS_GELU;
#pragma unroll 4
            for (int r = 0; r < 32; ++r) {
              const int row = q * 32 + r;
              const int hh = cty * CONV_TH + row / CONV_TW, ww = ctx * CONV_TW + row % CONV_TW;
              if (hh < p.cH && ww < p.cW) {
                float x = __fadd_rn(__uint_as_float(stg[r * 33 + lane]), b);
                if (gelu) x = gelu_erf(x);
                const int64_t o = (((int64_t)cb * p.cH + hh) * p.cW + ww) * p.N + n;
                if (p.out != nullptr) p.out[o] = x;
                if (p.out_pl[0] != nullptr) store_planes(p, o, x);
              }
            }
          } else if (n_ok && (p.epi == LVAE_EPI_SHUFFLE_NHWC || p.epi == LVAE_EPI_SHUFFLE_NCHW)) {
            // PixelShuffle(r) store: address = base(m) + offset(n), both separable (common.py:33-38)
            const float b = p.bias ? __ldg(p.bias + n) : 0.f;
            const int rr = p.r, Co = p.N / (rr * rr);
            const int qq = n / Co, cc = n - qq * Co, si = qq / rr, sj = qq - si * rr;
            const int Hr = p.Ho * rr, Wr = p.Wo * rr;
            const bool nhwc = p.epi == LVAE_EPI_SHUFFLE_NHWC;
            const int64_t coff = nhwc ? ((int64_t)si * Wr + sj) * Co + cc : ((int64_t)cc * Hr + si) * Wr + sj;
            int wo = row0 % p.Wo; int tq = row0 / p.Wo; int ho = tq % p.Ho; int bb = tq / p.Ho;
#pragma unroll 4
            for (int r = 0; r < 32; ++r) {
              if (row0 + r < p.M) {
                const int64_t base = nhwc ? (((int64_t)bb * Hr + ho * rr) * Wr + wo * rr) * Co
                                          : ((int64_t)bb * Co * Hr + ho * rr) * Wr + wo * rr;
                p.out[base + coff] = __fadd_rn(__uint_as_float(stg[r * 33 + lane]), b);
              }
              if (++wo == p.Wo) { wo = 0; if (++ho == p.Ho) { ho = 0; ++bb; } }
            }
          } else if (n_ok) {
#pragma unroll 2
            for (int r = 0; r < 32; ++r) {
              const int m = row0 + r;
              if (m < p.M) {
                if (p.atomic) atomicAdd(p.out + (int64_t)m * p.N + n, __uint_as_float(stg[r * 33 + lane]) * p.acc_scale);
                else epi_store(p, m, n, epi_value(p, m, n, __uint_as_float(stg[r * 33 + lane])));
              }
            }
          }
          __syncwarp();
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------- fp32 -> planes
// im2col + hi/lo split of an NHWC fp32 activation (same K order as the packed weights: (ky, kx, c) then segment 1)
__global__ void __launch_bounds__(256) split_im2col_kernel(
    const float* __restrict__ a0, const float* __restrict__ a1, int B, int H, int W, int Ho, int Wo,
    int C0, int C1, int ks, int stride, int pad, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
    __nv_bfloat16* __restrict__ l2, int64_t total4, int K, int f16, int act) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  pdl_trigger();
  pdl_wait();
  if (i >= total4) return;
  const int K4 = K >> 2;
  const int64_t m = i / K4; const int k = (int)(i - m * K4) * 4;
  const int K0 = ks * ks * C0;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (k < K0) {
    const int tap = k / C0, c = k - tap * C0, ky = tap / ks, kx = tap - ky * ks;
    const int wo = (int)(m % Wo); const int64_t t = m / Wo; const int ho = (int)(t % Ho); const int b = (int)(t / Ho);
    const int hh = ho * stride - pad + ky, ww = wo * stride - pad + kx;
    if (hh >= 0 && hh < H && ww >= 0 && ww < W)
      v = __ldg(reinterpret_cast<const float4*>(a0 + (((int64_t)b * H + hh) * W + ww) * C0 + c));
  } else {
    v = __ldg(reinterpret_cast<const float4*>(a1 + m * C1 + (k - K0)));
  }
  if (act) { v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w); }   // gelu(0) = 0: padding stays 0
  float2 a = make_float2(v.x, v.y), b = make_float2(v.z, v.w);
  const bool h16 = f16 != 0;
  uint2 w;
  w.x = split_next(a, h16); w.y = split_next(b, h16);
  *reinterpret_cast<uint2*>(hi + m * K + k) = w;
  if (lo) { w.x = split_next(a, h16); w.y = split_next(b, h16); *reinterpret_cast<uint2*>(lo + m * K + k) = w; }
  if (l2) { w.x = split_next(a, h16); w.y = split_next(b, h16); *reinterpret_cast<uint2*>(l2 + m * K + k) = w; }
}

// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

// [rows, K] bf16 row-major, box = [box_rows x 64], SWIZZLE_128B, zero fill outside the tensor
int make_map(CUtensorMap* map, const void* ptr, int64_t rows, int64_t K, int box_rows, int bk) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return LVAE_E_UNSUPPORTED; }
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (bk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) rows=%lld K=%lld box=%d", (int)r, (long long)rows, (long long)K, box_rows); return LVAE_E_BADARG; }
  return 0;
}

// NHWC bf16 plane viewed as (C, W, H, B); box = (bk channels, CONV_TW, CONV_TH, 1) = 128 rows of one swizzle row
static int make_map_conv(CUtensorMap* map, const void* ptr, int B, int H, int W, int C, int bk) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return LVAE_E_UNSUPPORTED; }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)bk, CONV_TW, CONV_TH, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (bk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (conv planes) failed (%d)", (int)r); return LVAE_E_BADARG; }
  return 0;
}

static void tc_geometry(const lvae_gemm_desc* d, int* Ho, int* Wo, int64_t* M, int* K) {
  *Ho = (d->H + 2 * d->pad - d->ksize) / d->stride + 1;
  *Wo = (d->W + 2 * d->pad - d->ksize) / d->stride + 1;
  *M = (int64_t)d->B * *Ho * *Wo;
  *K = d->ksize * d->ksize * d->C0 + ((d->a1 || d->a1_planes[0]) ? d->C1 : 0);
}

static int num_planes(int precision) {
  return precision == LVAE_PREC_BF16X6 ? 3 : ((precision == LVAE_PREC_BF16X3 || precision == LVAE_PREC_F16X3) ? 2 : 1);
}

int64_t gemm_tc_workspace_bytes(const lvae_gemm_desc* d) {
  if (d->a_planes[0]) return 0;
  int Ho, Wo, K; int64_t M;
  tc_geometry(d, &Ho, &Wo, &M, &K);
  return M * K * 2 * num_planes(d->precision);
}

// tuning knobs (lvae_set_tuning; initial values from the environment): [0] force BN (0 = automatic), [1] unused (was: fc1 planes stored
// straight from the row-owner registers -- 32 half-filled sectors per store instruction, 96 -> 137 us at H/8: dropped),
// [2] narrower N-tiles for GEMMs that do not fill the SMs
static int tc_tuning[4] = {-1, -1, -1, -1};
static int tc_tuning_get(int which) {
  if (tc_tuning[which] < 0) {
    static const char* names[4] = {"LVAE_TC_BN", "LVAE_TC_DIRECT", "LVAE_TC_AUTOBN", "LVAE_TC_RESERVED"};
    static const int defaults[4] = {0, 1, 1, 0};
    const char* e = getenv(names[which]);
    tc_tuning[which] = e ? atoi(e) : defaults[which];
  }
  return tc_tuning[which];
}
int gemm_tc_set_tuning(int which, int value) {
  if (which < 0 || which >= 4 || value < 0) return LVAE_E_BADARG;
  tc_tuning[which] = value;
  return 0;
}

// N-tile width of the 2-plane kernel when the default tiling leaves SMs idle or half a wave over: per 16-deep k-step an SM
// moves A (2 planes, TMA write + MMA read: 16 KB) plus 5 BN / 32 KB of B through its 128 B/clk shared-memory port
// (profiles/r2_mma_probe.md section 2), so a narrower tile is cheaper per step and there are more of them.  The result
// does not depend on BN: every output element is the same fixed-order sum over K (tests/test_gpu_kernels.py::
// test_gemm_tile_width_does_not_change_bits), so batch invariance holds although the choice looks at M.
static int pick_bn_fill(int M, int N, int K, int bn0, int n_sm) {
  const int mt = (M + TC_BM - 1) / TC_BM;
  int best_bn = bn0;
  double best = 1e30;
  for (int nt = (N + 127) / 128; nt <= (N + 31) / 32; ++nt) {
    int bn = (((N + nt - 1) / nt) + 15) & ~15;
    if (bn < 32 || bn > 128 || (N + bn - 1) / bn != nt) continue;
    const int waves = (mt * nt + n_sm - 1) / n_sm;
    const double cost = (double)waves * (K / 16.0) * (128.0 + 1.25 * bn) + 24.0 * bn;
    if (cost < best * 0.97) { best = cost; best_bn = bn; }     // ties and near-ties keep the wider tile
  }
  return best_bn;
}

static int pick_bn(int N, int npl) {
  const int maxbn = npl >= 2 ? 128 : 256;                      // split accumulators need 2 x BN columns per buffer
  const int nt = (N + maxbn - 1) / maxbn;
  int bn = (N + nt - 1) / nt;
  return (bn + 15) & ~15;
}

int gemm_tc_launch_split(const lvae_gemm_desc* d, int split_k, cudaStream_t stream);
static int gemm_tc_launch_impl(const lvae_gemm_desc* d, int split_k, const lvae_latent_epilogue* lat, cudaStream_t stream);
// gemm2_tc.cu: the CTA-pair (cta_group::2) kernel for the large plain GEMMs of the 2-plane modes
int gemm2_tc_launch(const lvae_gemm_desc* d, const void* const* a_pl, const void* const* a1_pl, int M, int K, int Ka, int C1,
                    int ek, cudaStream_t stream, int* handled);
int gemm_tc_launch(const lvae_gemm_desc* d, cudaStream_t stream) { return gemm_tc_launch_split(d, 0, stream); }

// split_k != 0: split-K over the persistent grid, partial tiles added atomically into d->out (which the caller zeroed);
// plain [M,K] x [N,K] only (LVAE_EPI_BIAS without bias)
int gemm_tc_launch_split(const lvae_gemm_desc* d, int split_k, cudaStream_t stream) { return gemm_tc_launch_impl(d, split_k, nullptr, stream); }
// lvae_gemm_latent: the implicit 3x3 convolution with the latent epilogue
int gemm_tc_launch_latent(const lvae_gemm_desc* d, const lvae_latent_epilogue* lat, cudaStream_t stream) {
  return gemm_tc_launch_impl(d, 0, lat, stream);
}
int gemm_tc_latent_num_partials(int H, int W, int N) {
  return ((H + CONV_TH - 1) / CONV_TH) * ((W + CONV_TW - 1) / CONV_TW) * ((N + 31) / 32) * 4;
}

// host-only: the N-tile width a plain [M, K] x [N, K] GEMM of this precision would be launched with on a device of n_sm SMs
// (lvae_gemm_tile_width; the launch itself asks the device for n_sm)
int gemm_tc_tile_width(int M, int N, int K, int precision, int n_sm) {
  const int npl = num_planes(precision);
  int bn = pick_bn(N, npl);
  const int mt0 = (M + TC_BM - 1) / TC_BM, nt0 = (N + bn - 1) / bn;
  if (npl == 2 && N >= 64 && mt0 * nt0 <= 2 * n_sm && tc_tuning_get(2)) bn = pick_bn_fill(M, N, K, bn, n_sm);
  return bn;
}

static int gemm_tc_launch_impl(const lvae_gemm_desc* d, int split_k, const lvae_latent_epilogue* lat, cudaStream_t stream) {
  const int npl = num_planes(d->precision);
  int Ho, Wo, K; int64_t M64;
  tc_geometry(d, &Ho, &Wo, &M64, &K);
  LVAE_CHECK_ARG(M64 > 0 && M64 < (1ll << 31));
  LVAE_CHECK_ARG(K % 8 == 0);                                  // 16-byte row pitch for the tensor maps
  for (int i = 0; i < npl; ++i) LVAE_CHECK_ARG(d->w_planes[i] != nullptr);
  const int M = (int)M64;
  const __nv_bfloat16* a_pl[3] = {(const __nv_bfloat16*)d->a_planes[0], (const __nv_bfloat16*)d->a_planes[1],
                                  (const __nv_bfloat16*)d->a_planes[2]};
  if (!a_pl[0]) {
    const int64_t need = gemm_tc_workspace_bytes(d);
    if (!d->workspace || d->workspace_bytes < need) {
      set_error("tensor-core GEMM from fp32 activations needs %lld workspace bytes, got %lld", (long long)need, (long long)d->workspace_bytes);
      return LVAE_E_BADARG;
    }
    __nv_bfloat16* ws = (__nv_bfloat16*)d->workspace;
    __nv_bfloat16* pl[3] = {ws, npl > 1 ? ws + (int64_t)M * K : nullptr, npl > 2 ? ws + 2 * (int64_t)M * K : nullptr};
    const int64_t total4 = (int64_t)M * (K / 4);
    LVAE_CUDA_CALL(launch_pdl(split_im2col_kernel, dim3((unsigned)((total4 + 255) / 256)), dim3(256), 0, stream,
        d->a0, d->a1, d->B, d->H, d->W, Ho, Wo, d->C0, d->a1 ? d->C1 : 0, d->ksize, d->stride, d->pad,
        pl[0], pl[1], pl[2], total4, K, (int)(d->precision == LVAE_PREC_F16X3 || d->precision == LVAE_PREC_F16), (int)d->a_act));
    LVAE_CUDA_LAUNCH_CHECK();
    for (int i = 0; i < 3; ++i) a_pl[i] = pl[i];
  } else {
    for (int i = 0; i < npl; ++i) LVAE_CHECK_ARG(a_pl[i] != nullptr);
  }

  // implicit 3x3 conv straight from NHWC planes (no im2col workspace): 9 shifted TMA box loads per channel block
  const bool conv = d->a_planes[0] != nullptr && d->ksize == 3 && d->stride == 1 && d->pad == 1 && d->a1 == nullptr &&
                    d->C0 % 16 == 0 && (d->epilogue == LVAE_EPI_BIAS || d->epilogue == LVAE_EPI_BIAS_GELU) &&
                    (d->out != nullptr || d->out_planes[0] != nullptr);
  if (d->a_planes[0] != nullptr && d->ksize != 1 && !conv) {
    set_error("pre-split A planes support 1x1 (plain [M,K]) and 3x3 stride-1 pad-1 convolutions with C %% 16 == 0 only");
    return LVAE_E_UNSUPPORTED;
  }
  const bool concat_planes = d->a_planes[0] != nullptr && d->a1_planes[0] != nullptr;
  if (concat_planes) {
    LVAE_CHECK_ARG(d->ksize == 1 && d->stride == 1 && d->pad == 0 && d->C0 % 64 == 0 && d->C1 > 0 && d->C1 % 8 == 0);
    for (int i = 0; i < npl; ++i) LVAE_CHECK_ARG(d->a1_planes[i] != nullptr);
  }
  TcParams p;
  p.M = M; p.N = d->N; p.K = K;
  p.k_split = concat_planes ? d->C0 : K;
  p.pl_act = d->out_planes_act;
  p.BN = pick_bn(d->N, npl);
  {
    static int sms = 0;
    if (sms == 0) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    const int mt0 = (M + TC_BM - 1) / TC_BM, nt0 = (d->N + p.BN - 1) / p.BN;
    if (npl == 2 && !conv && !split_k && d->N >= 64 && mt0 * nt0 <= 2 * sms && tc_tuning_get(2)) p.BN = pick_bn_fill(M, d->N, K, p.BN, sms);
    const int forced = tc_tuning_get(0);
    if (forced >= 16 && forced % 16 == 0 && forced <= (npl >= 2 ? 128 : 256)) p.BN = forced;
  }
  p.n_tiles = (d->N + p.BN - 1) / p.BN;
  p.conv = conv ? 1 : 0; p.cH = d->H; p.cW = d->W; p.cC = d->C0;
  p.tiles_w = conv ? (d->W + CONV_TW - 1) / CONV_TW : 1;
  p.tiles_h = conv ? (d->H + CONV_TH - 1) / CONV_TH : 1;
  p.num_tiles = (conv ? d->B * p.tiles_w * p.tiles_h : (M + TC_BM - 1) / TC_BM) * p.n_tiles;
  p.acc_cols = (npl >= 2 ? 2 : 1) * p.BN;
  int cols = 32; while (cols < 2 * p.acc_cols) cols <<= 1;
  p.tmem_cols = cols;
  p.out = d->out;
  for (int i = 0; i < 3; ++i) p.out_pl[i] = (i < npl) ? (__nv_bfloat16*)d->out_planes[i] : nullptr;
  if (p.out_pl[0] == nullptr) p.out_pl[1] = p.out_pl[2] = nullptr;
  if (p.out_pl[1] == nullptr) p.out_pl[2] = nullptr;
  // epilogue specialisation
  const bool shuffle = d->epilogue == LVAE_EPI_SHUFFLE_NHWC || d->epilogue == LVAE_EPI_SHUFFLE_NCHW;
  int ek = EK_MISC;
  if (conv || split_k) ek = EK_MISC;                           // tile rows are pixel patches: its own store loop | atomics
  else if (d->epilogue == LVAE_EPI_BIAS_GELU && p.out_pl[0] != nullptr && p.out == nullptr && d->N % 2 == 0) ek = EK_GELU;
  else if (!shuffle && d->N % 4 == 0) ek = EK_ROWS;
  if (npl == 2 && !conv && !split_k && (ek == EK_GELU || ek == EK_ROWS)) {
    // large plain GEMMs: 256 x BN tiles on CTA pairs (half the operand bytes per flop); bit-identical results
    const void* ap[2] = {a_pl[0], a_pl[1]};
    const void* a1p[2] = {d->a1_planes[0], d->a1_planes[1]};
    int handled = 0;
    const int rc2 = gemm2_tc_launch(d, ap, concat_planes ? a1p : nullptr, M, K, concat_planes ? d->C0 : K, d->C1,
                                    ek == EK_GELU ? 0 : 1, stream, &handled);
    if (rc2) return rc2;
    if (handled) return 0;
  }
  const int fixed = 1024 + tc_epi_stage_bytes(ek) + 256;       // alignment slack + transpose buffers + barriers
  const int budget = 227 * 1024 - fixed;
  // k-block: a 128-byte swizzle row (64 bf16) whenever two such stages fit (measured: 2 x 64 beats 4 x 32 on every
  // qarv shape), else 64-byte rows (32 bf16)
  p.BK = (2 * npl * (TC_BM + p.BN) * 64 * 2 <= budget) ? 64 : 32;
  if (K <= 32) p.BK = 32;
  if (conv) p.BK = d->C0 % 64 == 0 ? p.BK : (d->C0 % 32 == 0 ? 32 : 16);   // a k-block never straddles two filter taps
  { static const char* e = getenv("LVAE_TC_BK"); if (e) { const int v = atoi(e); if ((v == 64 || v == 32) && 2 * npl * (TC_BM + p.BN) * v * 2 <= budget) p.BK = v; } }   // tuning knob
  const int stage_bytes = npl * (TC_BM + p.BN) * p.BK * 2;
  int stages = budget / stage_bytes;
  if (stages > 8) stages = 8;
  const int nkb = (K + p.BK - 1) / p.BK;
  if (stages > nkb + 1) stages = nkb + 1;
  if (stages < 2) stages = 2;
  LVAE_CHECK_ARG(stages * stage_bytes <= budget);
  p.stages = stages;
  p.base_tiles = p.num_tiles; p.kb_per = nkb; p.atomic = 0;
  if (split_k) {
    LVAE_CHECK_ARG(!conv && !concat_planes && d->bias == nullptr && d->epilogue == LVAE_EPI_BIAS && d->out != nullptr &&
                   d->out_planes[0] == nullptr);
    int dev_sms = 148;
    { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, dev); }
    int splits = (2 * dev_sms + p.base_tiles - 1) / p.base_tiles;
    if (splits > nkb) splits = nkb;
    if (splits < 1) splits = 1;
    p.kb_per = (nkb + splits - 1) / splits;
    splits = (nkb + p.kb_per - 1) / p.kb_per;                  // every split owns at least one k-block
    p.num_tiles = p.base_tiles * splits;
    p.atomic = 1;
  }
  p.lat_prior = nullptr; p.lat_kl = nullptr; p.lat_kl_stride = 0; p.lat_kl_elem = nullptr; p.lat_sym = nullptr; p.lat_idx = nullptr;
  p.lat_table = nullptr; p.lat_nscales = 0; p.lat_cdf = 0;
  if (lat != nullptr) {
    if (!conv || ek != EK_MISC || p.n_tiles != 1 || d->epilogue != LVAE_EPI_BIAS || d->out == nullptr || d->out_planes[0] != nullptr) {
      set_error("lvae_gemm_latent needs the implicit 3x3 convolution (pre-split A planes, stride 1, pad 1, C %% 16 == 0), N <= 128, "
                "LVAE_EPI_BIAS and an fp32 output");
      return LVAE_E_UNSUPPORTED;
    }
    LVAE_CHECK_ARG(lat->prior != nullptr && lat->kl_partial != nullptr && lat->kl_stride >= gemm_tc_latent_num_partials(d->H, d->W, d->N));
    LVAE_CHECK_ARG((lat->sym == nullptr) == (lat->idx == nullptr));
    LVAE_CHECK_ARG(lat->sym == nullptr || (lat->scale_table != nullptr && lat->n_scales >= 1));
    p.lat_prior = lat->prior; p.lat_kl = lat->kl_partial; p.lat_kl_stride = lat->kl_stride; p.lat_kl_elem = lat->kl_elem;
    p.lat_sym = lat->sym; p.lat_idx = lat->idx; p.lat_table = lat->scale_table; p.lat_nscales = lat->n_scales; p.lat_cdf = lat->cdf_kind;
  }
  p.bias = d->bias; p.gamma = d->gamma; p.res = d->res;
  p.epi = d->epilogue; p.r = d->shuffle_r; p.Ho = Ho; p.Wo = Wo;
  { static const char* e = getenv("LVAE_TC_PREFETCH"); p.prefetch = e ? atoi(e) : 0; }
  p.f16 = (d->precision == LVAE_PREC_F16X3 || d->precision == LVAE_PREC_F16) ? 1 : 0;
  p.acc_scale = p.f16 ? 1.0f / LVAE_F16_WEIGHT_SCALE : 1.0f;
  LVAE_CHECK_ARG(p.out != nullptr || p.out_pl[0] != nullptr);

  TcMaps maps;
  int rc;
  for (int i = 0; i < 3; ++i) {
    const int j = i < npl ? i : 0;
    if (conv) { if ((rc = make_map_conv(&maps.a[i], a_pl[j], d->B, d->H, d->W, d->C0, p.BK))) return rc; }
    else if ((rc = make_map(&maps.a[i], a_pl[j], M, concat_planes ? d->C0 : K, TC_BM, p.BK))) return rc;
    if ((rc = make_map(&maps.b[i], d->w_planes[j], d->N, K, p.BN, p.BK))) return rc;
    if (concat_planes) { if ((rc = make_map(&maps.a1[i], d->a1_planes[j], M, d->C1, TC_BM, p.BK))) return rc; }
    else maps.a1[i] = maps.a[i];
  }

  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
#define LVAE_TC_ATTR(npl_, ek_) LVAE_CUDA_CALL(cudaFuncSetAttribute(gemm_tc_kernel<npl_, ek_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    LVAE_TC_ATTR(1, EK_GELU) LVAE_TC_ATTR(1, EK_ROWS) LVAE_TC_ATTR(1, EK_MISC)
    LVAE_TC_ATTR(2, EK_GELU) LVAE_TC_ATTR(2, EK_ROWS) LVAE_TC_ATTR(2, EK_MISC)
    LVAE_TC_ATTR(3, EK_GELU) LVAE_TC_ATTR(3, EK_ROWS) LVAE_TC_ATTR(3, EK_MISC)
#undef LVAE_TC_ATTR
  }
  const int smem = fixed + stages * stage_bytes;
  const int grid = p.num_tiles < n_sm ? p.num_tiles : n_sm;
#define LVAE_TC_LAUNCH(npl_, ek_) LVAE_CUDA_CALL(launch_pdl(gemm_tc_kernel<npl_, ek_>, dim3(grid), dim3(tc_threads(ek_)), (size_t)smem, stream, maps, p))
#define LVAE_TC_LAUNCH_EK(npl_) \
  do { if (ek == EK_GELU) LVAE_TC_LAUNCH(npl_, EK_GELU); else if (ek == EK_ROWS) LVAE_TC_LAUNCH(npl_, EK_ROWS); else LVAE_TC_LAUNCH(npl_, EK_MISC); } while (0)
  if (npl == 3) LVAE_TC_LAUNCH_EK(3);
  else if (npl == 2) LVAE_TC_LAUNCH_EK(2);
  else LVAE_TC_LAUNCH_EK(1);
#undef LVAE_TC_LAUNCH_EK
#undef LVAE_TC_LAUNCH
  LVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

}  // namespace lvae
