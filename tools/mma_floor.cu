// Microbenchmark (measurement aid): is the ~128-cycle cost of a tcgen05.mma.cta_group::2 (M = 256) with N < 256 a
// DEPENDENCY latency (same accumulator) or an issue-rate floor of the pipe? One thread issues 512 MMAs of K = 16 with
// operands resident in shared memory (contents irrelevant), round-robin over `nacc` accumulators.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I fcl_taco2_b200/csrc -o tools/_bin/mma_floor tools/mma_floor.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace fcl::umma;

__device__ __forceinline__ uint32_t cta_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* r, uint32_t n) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(r)), "r"(n) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t t, uint32_t n) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(t), "r"(n) : "memory");
}
__device__ __forceinline__ void mma2(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit2(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k(int n, int nacc, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  const int tid = threadIdx.x;
  for (int i = tid; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  cluster_sync();
  if (tid < 32) tmem_alloc2(&tbase, 512);
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  if (cta_rank() == 0 && tid == 0) {
    const uint32_t idesc = idesc_op_f32(256u, (uint32_t)n);
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 32 * 1024);
    const uint32_t b_lbo = (uint32_t)(n / 2) * 16u;
    uint64_t ad[4], bd[4];
    for (int i = 0; i < 4; ++i) {
      ad[i] = smem_desc(a_addr + (uint32_t)i * 4096u, 2048u, 128u);
      bd[i] = smem_desc(b_addr + (uint32_t)i * 2u * b_lbo, b_lbo, 128u);
    }
    uint32_t d[4];
    for (int i = 0; i < 4; ++i) d[i] = tbase + (uint32_t)((i % nacc) * n);
    const long long t0 = clock64();
    // first round overwrites, the rest accumulate; fully unrolled groups of 4 (the issue loop must not be the bound)
#pragma unroll 1
    for (int i = 0; i < 512; i += 4) {
      const uint32_t acc = i >= 4 ? 1u : 0u;
      mma2(d[0], ad[0], bd[0], idesc, acc);
      mma2(d[1], ad[1], bd[1], idesc, (nacc >= 2 || false) ? acc : 1u);
      mma2(d[2], ad[2], bd[2], idesc, nacc >= 4 ? acc : 1u);
      mma2(d[3], ad[3], bd[3], idesc, nacc >= 4 ? acc : 1u);
    }
    commit2(&bar);
    mbar_wait(&bar, 0);
    out[blockIdx.x / 2] = clock64() - t0;
  } else if (cta_rank() == 1 && tid == 0) {
    mbar_wait(&bar, 0);
  }
  tc_fence_before();
  cluster_sync();
  if (tid < 32) tmem_dealloc2(tbase, 512);
}

int main() {
  long long* out;
  cudaMalloc(&out, 74 * sizeof(long long));
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int n : {64, 128, 256})
    for (int nacc : {1, 2, 4}) {
      if (n * nacc > 512) continue;
      for (int rep = 0; rep < 2; ++rep) k<<<148, 128, 64 * 1024>>>(n, nacc, out);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[74];
      cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      double s = 0; for (int i = 0; i < 74; ++i) s += (double)h[i];
      printf("M=256 N=%3d, %d accumulator(s) round-robin: %.1f cycles per MMA (512 MMAs, 74 pairs)%s\n", n, nacc, s / 74 / 512,
             e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  return 0;
}
