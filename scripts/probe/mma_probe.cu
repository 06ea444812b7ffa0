// Cost of one tcgen05.mma.cta_group::1.kind::f16 (M = 128, K = 16) as a function of N, operands in shared memory
// (SWIZZLE_128B K-major tiles, contents irrelevant).  One CTA per SM; thread 0 issues `reps` back-to-back MMAs that
// accumulate into the same TMEM columns, commits, and waits.  Prints cycles per MMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe mma_probe.cu && ./mma_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__global__ void __launch_bounds__(128, 1) probe(int N, int reps, int nacc, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar; __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (128 * 64 * 2 + 256 * 64 * 2) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint64_t da = make_desc(smem_u32(smem)), db = make_desc(smem_u32(smem + 128 * 64 * 2));
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);     // fp16 in, fp32 out
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const uint64_t ko = (uint64_t)((r & 3) * 2);
      const uint32_t d = tmem + (uint32_t)((r % nacc) * N);
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(d), "l"(da + ko), "l"(db + ko), "r"(idesc), "r"(r >= nacc ? 1u : 0u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t done;
    do {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    } while (!done);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory"); }
}
int main() {
  long long* out; cudaMallocManaged(&out, 8);
  const int smem = 1024 + 128 * 64 * 2 + 256 * 64 * 2;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int Ns[] = {16, 32, 64, 96, 128, 192, 256};
  for (int nacc : {1, 2}) for (int N : Ns) {
    if (nacc * N > 512) continue;
    for (int grid : {1, 148}) {
      const int reps = 4096;
      probe<<<grid, 128, smem>>>(N, 64, nacc, out); cudaDeviceSynchronize();
      probe<<<grid, 128, smem>>>(N, reps, nacc, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("N=%d: %s\n", N, cudaGetErrorString(e)); return 1; }
      printf("M=128 N=%3d K=16, %d accumulator(s), grid %3d: %7.1f clk per MMA  (ideal at 8192 flop/clk/SM: %5.1f)\n", N, nacc, grid,
             (double)out[0] / reps, 2.0 * 128 * N * 16 / 8192.0);
    }
  }
  return 0;
}
