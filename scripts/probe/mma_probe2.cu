// tcgen05.mma cost probe, round 2: what does one kind::f16 MMA (K = 16) cost as a function of
//   * N (32 ... 256) at M = 128, operands in shared memory (SS), one issuing thread, one accumulator   [the r1 probe]
//   * the same with a tcgen05.commit every 4 MMAs, with operands rotating over 3 distinct shared-memory stages
//   * two issuing warps (different accumulators)
//   * M = 64
//   * the TS form (A operand from TMEM)
//   * cta_group::2 (M = 256 over a CTA pair, B split across the pair)
// Every variant is its own launch (so `ncu --metrics sm__pipe_tensor_cycles_active...` attributes per variant) on 148
// CTAs; CTA 0 reports clock64 cycles per MMA.  The cta_group::2 path is also CHECKED numerically against the host
// (exact small-integer operands), which pins the operand / accumulator layout the GEMM kernels rely on.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe2 mma_probe2.cu && ./mma_probe2
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// K-major SWIZZLE_128B tile, 64 16-bit elements (128 B) per row: SBO = 1024 B (8 rows), version 1, layout 2
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t idesc_f16(int M, int N) {   // fp16 x fp16 -> fp32, both K-major
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss1(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts1(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss2(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit1(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void commit2(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// element (row, k) of a [rows x 64] K-major SWIZZLE_128B tile: 16-byte chunk index XOR (row % 8)
__device__ __forceinline__ void put(uint8_t* tile, int row, int k, float v) {
  const int chunk = (k >> 3) ^ (row & 7);
  reinterpret_cast<__half*>(tile + row * 128 + chunk * 16)[k & 7] = __float2half(v);
}
__host__ __device__ inline float a_val(int r, int k) { return (float)(((r * 7 + k * 3) % 9) - 4); }
__host__ __device__ inline float b_val(int n, int k) { return (float)(((n * 5 + k * 11) % 7) - 3); }

constexpr int STAGE = 128 * 128 + 256 * 128;     // one A tile (128 x 64 fp16) + one B tile (256 x 64 fp16) = 48 KB

struct P1 { int M, N, reps, nacc, issuers, commit_every, rotate, ts; long long* out; };

__global__ void __launch_bounds__(128, 1) probe1(P1 p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar[2], dummy; __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nst = p.rotate ? 3 : 1;
  for (int i = threadIdx.x; i < nst * STAGE / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u ^ (uint32_t)(i * 2654435761u & 0x03ff03ffu);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&dummy)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (lane == 0 && warp < p.issuers) {
    const uint32_t idesc = idesc_f16(p.M, p.N);
    // issuer w accumulates into its own column range; the TS form keeps its A operand in columns 480..511
    const uint32_t dbase = tmem + (uint32_t)(warp * (p.issuers > 1 ? 256 : 0));
    const long long t0 = clock64();
    for (int r = 0; r < p.reps; ++r) {
      const uint32_t st = smem_u32(smem) + (uint32_t)((p.rotate ? (r >> 2) % 3 : 0) * STAGE);
      const uint64_t ko = (uint64_t)((r & 3) * 2);
      const uint64_t da = make_desc(st) + ko, db = make_desc(st + 128 * 128) + ko;
      const uint32_t d = dbase + (uint32_t)((r % p.nacc) * p.N);
      if (p.ts) mma_ts1(d, tmem + 480u, db, idesc, r >= p.nacc ? 1u : 0u);
      else mma_ss1(d, da, db, idesc, r >= p.nacc ? 1u : 0u);
      if (p.commit_every && (r % p.commit_every) == p.commit_every - 1) commit1(smem_u32(&dummy));
    }
    commit1(smem_u32(&bar[warp]));
    mbar_wait(smem_u32(&bar[warp]), 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0 && warp == 0) p.out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory"); }
}

// cta_group::2: the pair computes D[256 x N] = A[256 x K] * B[N x K]^T; CTA r holds A rows [128 r, 128 r + 128), B rows
// [N/2 r, N/2 r + N/2) and receives D rows [128 r, 128 r + 128) x all N columns in its own TMEM.
struct P2 { int N, reps, ksteps, check; long long* out; float* D; };

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe2(P2 p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar; __shared__ uint32_t slot;
  uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* At = smem; uint8_t* Bt = smem + 128 * 128;
  for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) put(At, i >> 6, i & 63, a_val((int)rank * 128 + (i >> 6), i & 63));
  for (int i = threadIdx.x; i < (p.N / 2) * 64; i += blockDim.x) put(Bt, i >> 6, i & 63, b_val((int)rank * (p.N / 2) + (i >> 6), i & 63));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    long long t0 = 0;
    if (rank == 0) {
      const uint32_t idesc = idesc_f16(256, p.N);
      const uint64_t da = make_desc(smem_u32(At)), db = make_desc(smem_u32(Bt));
      t0 = clock64();
      for (int r = 0; r < p.reps; ++r) {
        const uint64_t ko = (uint64_t)((r % p.ksteps) * 2);
        mma_ss2(tmem, da + ko, db + ko, idesc, r > 0 ? 1u : 0u);
      }
      commit2(smem_u32(&bar), (uint16_t)3);      // arrives on `bar` of both CTAs once the MMAs have completed
    }
    mbar_wait(smem_u32(&bar), 0);
    if (rank == 0 && blockIdx.x == 0) p.out[0] = clock64() - t0;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (p.check && blockIdx.x < 2) {
    for (int c = 0; c < p.N; c += 32) {
      uint32_t v[32];
      ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
      const int row = (int)rank * 128 + warp * 32 + lane;
      for (int j = 0; j < 32; ++j) p.D[(size_t)row * p.N + c + j] = __uint_as_float(v[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 0) { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory"); }
}

static long long* g_out;
static double run1(int M, int N, int nacc, int issuers, int commit_every, int rotate, int ts, int reps = 4096) {
  const int smem = 1024 + 3 * STAGE;
  P1 p{M, N, 64, nacc, issuers, commit_every, rotate, ts, g_out};
  probe1<<<148, 128, smem>>>(p); cudaDeviceSynchronize();
  p.reps = reps;
  probe1<<<148, 128, smem>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("probe1 M=%d N=%d: %s\n", M, N, cudaGetErrorString(e)); exit(1); }
  return (double)g_out[0] / reps;
}

int main() {
  cudaMallocManaged(&g_out, 8);
  cudaFuncSetAttribute(probe1, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 + 3 * STAGE);
  cudaFuncSetAttribute(probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 + STAGE);
  const int Ns[] = {32, 64, 128, 192, 256};
  printf("# clk per tcgen05.mma kind::f16 K=16 (148 CTAs, CTA 0 reports); ideal = M*N/256 at 8192 flop/clk/SM\n");
  printf("%-44s", "variant \\ N"); for (int N : Ns) printf("%8d", N); printf("\n");
  auto row = [&](const char* name, int M, int nacc, int issuers, int ce, int rot, int ts) {
    printf("%-44s", name);
    for (int N : Ns) {
      if ((issuers > 1 ? 2 : 1) * nacc * N > 480) { printf("%8s", "-"); continue; }
      printf("%8.1f", run1(M, N, nacc, issuers, ce, rot, ts));
    }
    printf("\n"); fflush(stdout);
  };
  row("cg1 M=128 SS, 1 issuer, 1 acc (r1 probe)", 128, 1, 1, 0, 0, 0);
  row("cg1 M=128 SS, commit every 4", 128, 1, 1, 4, 0, 0);
  row("cg1 M=128 SS, commit/4, 3 rotating stages", 128, 1, 1, 4, 1, 0);
  row("cg1 M=128 SS, 2 accumulators alternating", 128, 2, 1, 4, 1, 0);
  row("cg1 M=128 SS, 2 issuing warps (per issuer)", 128, 1, 2, 4, 1, 0);
  row("cg1 M=64  SS", 64, 1, 1, 4, 1, 0);
  row("cg1 M=128 TS (A from TMEM)", 128, 1, 1, 4, 1, 1);
  printf("%-44s", "ideal M=128"); for (int N : Ns) printf("%8.1f", 128.0 * N / 256); printf("\n");

  // cta_group::2: timing, then the numerical check
  printf("%-44s", "cg2 M=256 SS (per pair instruction)");
  for (int N : Ns) {
    P2 p{N, 64, 4, 0, g_out, nullptr};
    probe2<<<148, 128, 1024 + STAGE>>>(p); cudaDeviceSynchronize();
    p.reps = 4096;
    probe2<<<148, 128, 1024 + STAGE>>>(p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("probe2 N=%d: %s\n", N, cudaGetErrorString(e)); return 1; }
    printf("%8.1f", (double)g_out[0] / 4096);
  }
  printf("\n%-44s", "ideal cg2 (M=256: N/2 per SM)"); for (int N : Ns) printf("%8.1f", N / 2.0); printf("\n");
  for (int N : {64, 256}) {
    float* D; cudaMallocManaged(&D, (size_t)256 * N * 4);
    for (int i = 0; i < 256 * N; ++i) D[i] = -12345.f;
    P2 p{N, 4, 4, 1, g_out, D};               // 4 k-steps = the whole K = 64 tile, once
    probe2<<<2, 128, 1024 + STAGE>>>(p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("probe2 check N=%d: %s\n", N, cudaGetErrorString(e)); return 1; }
    int bad = 0; double maxd = 0;
    for (int r = 0; r < 256; ++r) for (int n = 0; n < N; ++n) {
      float ref = 0; for (int k = 0; k < 64; ++k) ref += a_val(r, k) * b_val(n, k);
      const double dd = fabs((double)ref - D[(size_t)r * N + n]);
      if (dd > maxd) maxd = dd;
      if (dd != 0 && bad++ < 5) printf("  mismatch r=%d n=%d got %g want %g\n", r, n, D[(size_t)r * N + n], ref);
    }
    printf("cta_group::2 check N=%d: %d mismatches of %d, max |diff| %g  => D rows [128 r, +128) in CTA r's TMEM, columns [N/2 r', +N/2) from CTA r' B rows: %s\n",
           N, bad, 256 * N, maxd, bad ? "NOT CONFIRMED" : "confirmed");
    cudaFree(D);
  }
  return 0;
}
