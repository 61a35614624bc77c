// Microbenchmark (measurement aid, not product code): what bounds a tcgen05.mma stream on B200?
//   * intrinsic issue rate of M=128 (cta_group::1) and M=256 (cta_group::2) bf16 MMAs, N=256, per K=64 "stage",
//     with un-swizzled (SWIZZLE_NONE, the layout the kernels in csrc/ use) and SWIZZLE_128B descriptors,
//     operands resident in shared memory (contents are irrelevant for timing);
//   * the per-SM ingest rate of cp.async.bulk from L2 (no MMA);
//   * both at once, un-synchronised: does the bulk-copy stream slow the MMA stream down (and vice versa)?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I fcl_taco2_b200/csrc -o tools/_bin/umma_rate tools/umma_rate.cu
// Run:   tools/_bin/umma_rate          (prints one line per configuration)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "umma.cuh"

using namespace fcl::umma;

struct Cfg {
  int pair;          // 0: cta_group::1, M=128; 1: cta_group::2, M=256 (2-CTA cluster)
  int swizzle;       // 0: SWIZZLE_NONE core-matrix image; 1: SWIZZLE_128B
  int mma_stages;    // K=64 stages issued by the MMA thread (0 = no MMA)
  int load_bytes;    // bytes per bulk copy of the producer warp (0 = no copies)
  int n_loads;       // copies issued by the producer warp
  int a_stage_bytes; // operand ring geometry the MMA descriptors walk
  int b_stage_bytes;
  int n_op_stages;
  int noise;         // extra 16 warps: 0 none, 1 tcgen05.ld loop, 2 integer multiply loop, 3 MUFU loop, 4 global ld/st, 5 mbarrier polling
  int noise_first, noise_count;   // which warps run the noise loop (warp w sits on scheduler w % 4; the MMA thread is in warp 1)
  int commit_every;  // stages per tcgen05.commit (1 = what a ring with per-stage release does); 16 stages form a wait group
};

struct Out { long long mma_clk, load_clk; };

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
// K-major SWIZZLE_128B descriptor: rows of 64 bf16 (128 B), 8-row atoms 1024 B apart; layout type 2 in bits 61-63
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

template <bool PAIR>
__global__ void __launch_bounds__(640, 1) rate_kernel(Cfg c, const uint8_t* __restrict__ src, size_t src_bytes, Out* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // swizzle atoms are 1024-byte aligned
  __shared__ uint64_t done[2], ld[4];
  __shared__ uint32_t tmem_base;
  __shared__ volatile int stop;
  __shared__ uint64_t poll_bar;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = PAIR ? cta_rank() : 0u;
  if (threadIdx.x == 0) {
    stop = 0;
    mbar_init(&poll_bar, 1);
    mbar_init(&done[0], 16 / c.commit_every); mbar_init(&done[1], 16 / c.commit_every);
    for (int i = 0; i < 4; ++i) mbar_init(&ld[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if constexpr (PAIR) cluster_sync();
  if (warp == 2) { if constexpr (PAIR) tmem_alloc2(&tmem_base, 512); else tmem_alloc(&tmem_base, 512); }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync();
  tc_fence_after();
  const uint32_t tmem = tmem_base;
  const uint32_t op_bytes = (uint32_t)(c.a_stage_bytes + c.b_stage_bytes) * c.n_op_stages;
  uint8_t* land = smem + op_bytes;                       // landing zone of the producer's copies (4 slots)

  if (warp == 1 && c.mma_stages > 0 && rank == 0) {
    if (elect_one()) {
      const uint32_t idesc = idesc_op_f32(PAIR ? 256u : 128u, 256u);
      const uint32_t b_rows = PAIR ? 128u : 256u;      // rows of B held by THIS CTA
      const long long t0 = clock64();
      const int group = 16;                              // commit every 16 stages, wait one group behind
      int ngroups = 0;
      for (int s = 0; s < c.mma_stages; ++s) {
        const int slot = s % c.n_op_stages;
        const uint32_t a0 = smem_u32(smem + (size_t)slot * (c.a_stage_bytes + c.b_stage_bytes));
        const uint32_t b0 = a0 + c.a_stage_bytes;
        const uint32_t d = tmem + (uint32_t)((s / 12) & 1) * 256u;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint64_t ad, bd;
          if (c.swizzle) { ad = desc_sw128(a0 + k * 32); bd = desc_sw128(b0 + k * 32); }
          else {
            ad = smem_desc(a0 + k * 2 * 128 * 16, 128 * 16, 128);
            bd = smem_desc(b0 + k * 2 * b_rows * 16, b_rows * 16, 128);
          }
          if constexpr (PAIR) mma2(d, ad, bd, idesc, (s % 12) | k); else mma_bf16_ss(d, ad, bd, idesc, (s % 12) | k);
        }
        if (((s + 1) & (c.commit_every - 1)) == 0) {
          if constexpr (PAIR) commit2(&done[ngroups & 1]); else mma_commit(&done[ngroups & 1]);
        }
        if ((s + 1) % group == 0) {
          if (ngroups >= 1) mbar_wait(&done[(ngroups - 1) & 1], (uint32_t)(((ngroups - 1) >> 1) & 1));
          ++ngroups;
        }
      }
      mbar_wait(&done[(ngroups - 1) & 1], (uint32_t)(((ngroups - 1) >> 1) & 1));
      out[blockIdx.x].mma_clk = clock64() - t0;
      stop = 1;
    }
    __syncwarp();
  }
  if (PAIR && rank == 1 && warp == 1) stop = 1;          // (the peer's noise warps run a fixed, short count instead)
  if (warp == 0 && c.n_loads > 0) {
    if (elect_one()) {
      // every CTA walks its own window of the (L2-resident) source so that the copies are not all the same lines
      size_t off = ((size_t)blockIdx.x * 1315423911ull) % (src_bytes / 2) / 1024 * 1024;
      const long long t0 = clock64();
      for (int i = 0; i < c.n_loads; ++i) {
        const int slot = i & 3;
        if (i >= 4) mbar_wait(&ld[slot], (uint32_t)(((i >> 2) - 1) & 1));
        mbar_arrive_expect_tx(&ld[slot], (uint32_t)c.load_bytes);
        bulk_g2s(land + (size_t)slot * c.load_bytes, src + off, (uint32_t)c.load_bytes, &ld[slot]);
        off += c.load_bytes;
        if (off + c.load_bytes > src_bytes) off = 0;
      }
      for (int i = (c.n_loads > 4 ? c.n_loads - 4 : 0); i < c.n_loads; ++i) mbar_wait(&ld[i & 3], (uint32_t)((i >> 2) & 1));
      out[blockIdx.x].load_clk = clock64() - t0;
    }
    __syncwarp();
  }
  if (warp >= c.noise_first && warp < c.noise_first + c.noise_count && c.noise) {
    // background work of the kind the decoder's epilogue warps do, until the MMA thread of this CTA (pair: the leader's) is done
    const int q = warp & 3, lane = threadIdx.x & 31;
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
    float acc = 0.f; uint32_t x = threadIdx.x * 2654435761u + 1u;
    float* g = reinterpret_cast<float*>(const_cast<uint8_t*>(src)) + ((size_t)blockIdx.x * 512 + (threadIdx.x & 511)) * 64;
    for (int it = 0; it < 100000 && !stop; ++it) {
      if (c.noise == 1) { float v[16]; tmem_ld16(lane_addr + (uint32_t)((it & 31) * 16), v); acc += v[0] + v[15]; }
      else if (c.noise == 2) {
#pragma unroll
        for (int k = 0; k < 32; ++k) { const uint64_t m64 = (uint64_t)x * 0xD2511F53u; x = (uint32_t)(m64 >> 32) ^ (uint32_t)m64 ^ 0x9E3779B9u; }
      } else if (c.noise == 3) {
#pragma unroll
        for (int k = 0; k < 16; ++k) acc = tanh_fast(acc + 0.37f);
      } else if (c.noise == 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { const float4 t4 = __ldcg(reinterpret_cast<const float4*>(g) + ((it + k) & 15)); acc += t4.x; }
        reinterpret_cast<float4*>(g)[it & 15] = make_float4(acc, 0.f, 0.f, 0.f);
      } else if (c.noise == 5) {
        if (mbar_try_wait(&poll_bar, 0)) acc += 1.f;    // never completes: every call is a (possibly suspended) poll
      }
    }
    if (acc == 123.456f || x == 77u) out[blockIdx.x].load_clk = 1;   // keep the work alive
    (void)lane;
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync();
  if (warp == 2) { if constexpr (PAIR) tmem_dealloc2(tmem, 512); else tmem_dealloc(tmem, 512); }
}


// ---------------------------------------------------------------------------------------------------------------
// Synchronised streaming GEMM skeleton, the handshake of csrc/decoder_bf16{,_pair}.cu without the epilogue:
// producer warp -> ring (full/empty mbarriers) -> MMA thread; pair mode adds the peer's "my half landed" relay.
constexpr int kMaxStages = 8;
struct StreamSh {
  uint64_t full[kMaxStages], empty[kMaxStages], peer_full[kMaxStages], done;
  uint32_t tmem_base;
};
__device__ __forceinline__ void arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
               "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}

__device__ __forceinline__ void mbar_spin(uint64_t* bar, uint32_t parity) {   // non-blocking test_wait poll
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
}

__device__ __forceinline__ void arrive_remote_relaxed(uint64_t* bar, uint32_t cta) {
  asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
               "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}

template <bool PAIR>
__global__ void __launch_bounds__(128, 1) stream_kernel(int n_stages, int ring, int a_bytes, int b_bytes, int split_copies,
                                                        const uint8_t* __restrict__ src, size_t src_bytes, Out* out, unsigned long long* tr) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ StreamSh sh;
  // timeline of stages 600..615 of cluster 0 in ns (%globaltimer is common to all SMs): tr[(s - 600) * 8 + event]
#define TR(ev) do { if (tr && blockIdx.x < 2 && s >= 600 && s < 616) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); tr[(s - 600) * 8 + (ev)] = t_; } } while (0)
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = PAIR ? cta_rank() : 0u;
  const uint32_t stage_bytes = (uint32_t)(a_bytes + b_bytes);
  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxStages; ++i) { mbar_init(&sh.full[i], 1); mbar_init(&sh.empty[i], 1); mbar_init(&sh.peer_full[i], 1); }
    mbar_init(&sh.done, 1);
    fence_barrier_init();
  }
  __syncthreads();
  if constexpr (PAIR) cluster_sync();
  if (warp == 2) { if constexpr (PAIR) tmem_alloc2(&sh.tmem_base, 512); else tmem_alloc(&sh.tmem_base, 512); }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync();
  tc_fence_after();
  const uint32_t tmem = sh.tmem_base;
  if (warp == 0) {
    if (elect_one()) {
      size_t off = ((size_t)blockIdx.x * 1315423911ull) % (src_bytes / 2) / 1024 * 1024;
      uint32_t stage = 0, ph = 0;
      for (int s = 0; s < n_stages; ++s) {
        if (split_copies & 2) mbar_spin(&sh.empty[stage], ph ^ 1u); else mbar_wait(&sh.empty[stage], ph ^ 1u);
        TR(rank == 0 ? 0 : 1);                           // producer: slot free seen
        mbar_arrive_expect_tx(&sh.full[stage], stage_bytes);
        uint8_t* dst = smem + (size_t)stage * stage_bytes;
        if (split_copies & 1) {
          bulk_g2s(dst, src + off, (uint32_t)a_bytes, &sh.full[stage]);
          bulk_g2s(dst + a_bytes, src + off + (4u << 20), (uint32_t)b_bytes, &sh.full[stage]);
        } else {
          bulk_g2s(dst, src + off, stage_bytes, &sh.full[stage]);
        }
        off += stage_bytes;
        if (off + stage_bytes + (4u << 20) > src_bytes) off = 0;
        if (++stage == (uint32_t)ring) { stage = 0; ph ^= 1u; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = idesc_op_f32(PAIR ? 256u : 128u, 256u);
      const uint32_t b_rows = (uint32_t)b_bytes / 128u;
      uint32_t stage = 0, ph = 0;
      const long long t0 = clock64();
      for (int s = 0; s < n_stages; ++s) {
        mbar_wait(&sh.full[stage], ph);
        TR(rank == 0 ? 2 : 3);                           // own copy landed
        if (PAIR && rank == 1) {
          if (!(split_copies & 8)) { if (split_copies & 4) arrive_remote_relaxed(&sh.peer_full[stage], 0); else arrive_remote(&sh.peer_full[stage], 0); }
        } else {
          if constexpr (PAIR) { if (split_copies & 2) mbar_spin(&sh.peer_full[stage], ph); else mbar_wait(&sh.peer_full[stage], ph); }
          TR(4);                                         // leader: peer's half landed
          tc_fence_after();
          const uint32_t a0 = smem_u32(smem + (size_t)stage * stage_bytes), b0 = a0 + a_bytes;
          const uint32_t d = tmem + (uint32_t)((s / 12) & 1) * 256u;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = smem_desc(a0 + k * 2 * 128 * 16, 128 * 16, 128);
            const uint64_t bd = smem_desc(b0 + k * 2 * b_rows * 16, b_rows * 16, 128);
            if constexpr (PAIR) mma2(d, ad, bd, idesc, (s % 12) | k); else mma_bf16_ss(d, ad, bd, idesc, (s % 12) | k);
          }
          if constexpr (PAIR) commit2(&sh.empty[stage]); else mma_commit(&sh.empty[stage]);
          TR(5);                                         // leader: MMAs + commit issued
        }
        if (++stage == (uint32_t)ring) { stage = 0; ph ^= 1u; }
      }
      if (rank == 0) {
        if constexpr (PAIR) commit2(&sh.done); else mma_commit(&sh.done);
        mbar_wait(&sh.done, 0);
        out[blockIdx.x].mma_clk = clock64() - t0;
      }
    }
    __syncwarp();
  }
  if (PAIR && rank == 1 && warp == 3 && (split_copies & 8) && (threadIdx.x & 31) < ring) {
    // one relay lane per ring slot: the remote arrivals of different slots overlap
    const int slot = threadIdx.x & 31;
    uint32_t par = 0;
    for (int s = slot; s < n_stages; s += ring) {
      mbar_wait(&sh.full[slot], par);
      if (split_copies & 4) arrive_remote_relaxed(&sh.peer_full[slot], 0); else arrive_remote(&sh.peer_full[slot], 0);
      par ^= 1u;
    }
  }
  if (PAIR && rank == 1 && threadIdx.x == 64) mbar_wait(&sh.done, 0);   // the peer gets the multicast "all done" too
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync();
  if (warp == 2) { if constexpr (PAIR) tmem_dealloc2(tmem, 512); else tmem_dealloc(tmem, 512); }
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

static void run(const char* name, Cfg c, int n_ctas, const uint8_t* src, size_t src_bytes, Out* d_out) {
  const size_t smem = (size_t)(c.a_stage_bytes + c.b_stage_bytes) * c.n_op_stages + (size_t)4 * c.load_bytes + 2048;
  auto kern = c.pair ? rate_kernel<true> : rate_kernel<false>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaMemset(d_out, 0, sizeof(Out) * 148));
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(n_ctas); lc.blockDim = dim3(c.noise ? 640 : 128); lc.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = c.pair ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  lc.attrs = at; lc.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaLaunchKernelEx(&lc, kern, c, src, src_bytes, d_out));
    CK(cudaDeviceSynchronize());
  }
  std::vector<Out> h(148);
  CK(cudaMemcpy(h.data(), d_out, sizeof(Out) * 148, cudaMemcpyDeviceToHost));
  double mma = 0, ldc = 0; int nm = 0, nl = 0;
  for (int i = 0; i < n_ctas; ++i) {
    if (h[i].mma_clk) { mma += (double)h[i].mma_clk; ++nm; }
    if (h[i].load_clk) { ldc += (double)h[i].load_clk; ++nl; }
  }
  printf("%-44s ctas %3d", name, n_ctas);
  if (nm) printf("  mma: %7.1f clk per K=64 stage", mma / nm / c.mma_stages);
  if (nl) printf("  bulk copy: %6.1f B/clk per SM (%d B copies)", (double)c.load_bytes * c.n_loads / (ldc / nl), c.load_bytes);
  printf("\n");
}


static unsigned long long* g_trace = nullptr;
static void run_stream(const char* name, bool pair, int ring, int a_bytes, int b_bytes, int split, int n_ctas,
                       const uint8_t* src, size_t src_bytes, Out* d_out) {
  const int n_stages = 1200;
  const size_t smem = (size_t)(a_bytes + b_bytes) * ring + 2048;
  auto kern = pair ? stream_kernel<true> : stream_kernel<false>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaMemset(d_out, 0, sizeof(Out) * 148));
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(n_ctas); lc.blockDim = dim3(128); lc.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = pair ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  lc.attrs = at; lc.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaLaunchKernelEx(&lc, kern, n_stages, ring, a_bytes, b_bytes, split, src, src_bytes, d_out, g_trace));
    CK(cudaDeviceSynchronize());
  }
  std::vector<Out> h(148);
  CK(cudaMemcpy(h.data(), d_out, sizeof(Out) * 148, cudaMemcpyDeviceToHost));
  double mma = 0; int nm = 0;
  for (int i = 0; i < n_ctas; ++i) if (h[i].mma_clk) { mma += (double)h[i].mma_clk; ++nm; }
  if (g_trace && getenv("UMMA_TRACE")) {
    std::vector<unsigned long long> t(16 * 8);
    CK(cudaMemcpy(t.data(), g_trace, sizeof(unsigned long long) * 16 * 8, cudaMemcpyDeviceToHost));
    printf("  stage: prod0 free | prod1 free | landed0 | landed1 | leader sees peer | issued   (ns since stage 600's first event)\n");
    unsigned long long t0 = ~0ull;
    for (int e = 0; e < 6; ++e) if (t[e] && t[e] < t0) t0 = t[e];
    for (int i = 0; i < 16; ++i) {
      printf("  %3d:", 600 + i);
      for (int e = 0; e < 6; ++e) printf(" %8lld", t[i * 8 + e] ? (long long)(t[i * 8 + e] - t0) : -1ll);
      printf("\n");
    }
    CK(cudaMemset(g_trace, 0, sizeof(unsigned long long) * 16 * 8));
  }
  printf("%-52s ctas %3d  ring %d x %2d KB  %7.1f clk per K=64 stage  (%5.1f B/clk per SM)\n", name, n_ctas, ring,
         (a_bytes + b_bytes) >> 10, mma / nm / n_stages, (double)(a_bytes + b_bytes) * n_stages / (mma / nm));
}

int main(int argc, char** argv) {
  const bool swz = !(argc > 1 && argv[1][0] == 'n');     // `umma_rate n`: skip the SWIZZLE_128B configurations
  const size_t src_bytes = 32u << 20;                    // L2-resident source
  uint8_t* src; Out* d_out;
  CK(cudaMalloc(&src, src_bytes)); CK(cudaMemset(src, 0, src_bytes));
  CK(cudaMalloc(&d_out, sizeof(Out) * 148));
  const int S = 1536;
  CK(cudaMalloc(&g_trace, sizeof(unsigned long long) * 16 * 8)); CK(cudaMemset(g_trace, 0, sizeof(unsigned long long) * 16 * 8));
  for (int n : {148, 16}) {
    // MMA only
    run("cta1 M128 N256 no-swizzle, MMA only", Cfg{0, 0, S, 0, 0, 16384, 32768, 3, 0, 4, 16, 16}, n, src, src_bytes, d_out);
    if (swz) run("cta1 M128 N256 swizzle128, MMA only", Cfg{0, 1, S, 0, 0, 16384, 32768, 3, 0, 4, 16, 16}, n, src, src_bytes, d_out);
    run("cta2 M256 N256 no-swizzle, MMA only", Cfg{1, 0, S, 0, 0, 16384, 16384, 4, 0, 4, 16, 16}, n, src, src_bytes, d_out);
    if (swz) run("cta2 M256 N256 swizzle128, MMA only", Cfg{1, 1, S, 0, 0, 16384, 16384, 4, 0, 4, 16, 16}, n, src, src_bytes, d_out);
    for (int ce : {1, 2, 4}) {
      char nm[64];
      snprintf(nm, sizeof nm, "cta1 no-swizzle, MMA only, commit every %d", ce);
      run(nm, Cfg{0, 0, S, 0, 0, 16384, 32768, 3, 0, 4, 16, ce}, n, src, src_bytes, d_out);
      snprintf(nm, sizeof nm, "cta2 no-swizzle, MMA only, commit every %d", ce);
      run(nm, Cfg{1, 0, S, 0, 0, 16384, 16384, 4, 0, 4, 16, ce}, n, src, src_bytes, d_out);
    }
    for (int nz = 1; nz <= 5; ++nz) {
      const char* what[] = {"", "tcgen05.ld", "integer multiply", "MUFU tanh", "global ld/st", "mbarrier try_wait polling"};
      char nm[80];
      snprintf(nm, sizeof nm, "cta1 MMA only + 16 warps of %s", what[nz]);
      run(nm, Cfg{0, 0, S, 0, 0, 16384, 32768, 3, nz, 4, 16, 16}, n, src, src_bytes, d_out);
    }
    // where does the interference come from? noise warps that do NOT share the MMA thread's scheduler (warp 1 -> scheduler 1)
    run("cta1 MMA only + integer multiply in warps 2,3", Cfg{0, 0, S, 0, 0, 16384, 32768, 3, 2, 2, 2, 16}, n, src, src_bytes, d_out);
    run("cta1 MMA only + integer multiply in warp 5 (same scheduler)", Cfg{0, 0, S, 0, 0, 16384, 32768, 3, 2, 5, 1, 16}, n, src, src_bytes, d_out);
    run("cta1 MMA only + integer multiply in warps 6,7,8", Cfg{0, 0, S, 0, 0, 16384, 32768, 3, 2, 6, 3, 16}, n, src, src_bytes, d_out);
    run("cta1 MMA only + MUFU in warps 6,7,8", Cfg{0, 0, S, 0, 0, 16384, 32768, 3, 3, 6, 3, 16}, n, src, src_bytes, d_out);
    run("cta1 MMA only + tcgen05.ld in warps 6,7,8", Cfg{0, 0, S, 0, 0, 16384, 32768, 3, 1, 6, 3, 16}, n, src, src_bytes, d_out);
    // copies only
    for (int lb : {8192, 16384, 32768})
      run("bulk copies only", Cfg{0, 0, 0, lb, 4096, 16384, 32768, 1, 0, 4, 16, 16}, n, src, src_bytes, d_out);
    // both, un-synchronised
    run("cta1 no-swizzle + 16 KB copies", Cfg{0, 0, S, 16384, 6000, 16384, 32768, 3, 0, 4, 16, 16}, n, src, src_bytes, d_out);
    if (swz) run("cta1 swizzle128 + 16 KB copies", Cfg{0, 1, S, 16384, 6000, 16384, 32768, 3, 0, 4, 16, 16}, n, src, src_bytes, d_out);
    run("cta2 no-swizzle + 16 KB copies", Cfg{1, 0, S, 16384, 6000, 16384, 16384, 4, 0, 4, 16, 16}, n, src, src_bytes, d_out);
    if (swz) run("cta2 swizzle128 + 16 KB copies", Cfg{1, 1, S, 16384, 6000, 16384, 16384, 4, 0, 4, 16, 16}, n, src, src_bytes, d_out);
  }
  for (int n : {148, 16}) {
    run_stream("stream cta1 A16+B32, one copy per stage", false, 4, 16384, 32768, 0, n, src, src_bytes, d_out);
    run_stream("stream cta1 A16+B32, two copies per stage", false, 4, 16384, 32768, 1, n, src, src_bytes, d_out);
    run_stream("stream cta1 B32 only (A resident)", false, 6, 0, 32768, 0, n, src, src_bytes, d_out);
    run_stream("stream cta2 A16+B16, one copy per stage", true, 6, 16384, 16384, 0, n, src, src_bytes, d_out);
    run_stream("stream cta2 A16+B16, two copies per stage", true, 6, 16384, 16384, 1, n, src, src_bytes, d_out);
    run_stream("stream cta2 B16 only (A resident)", true, 8, 0, 16384, 0, n, src, src_bytes, d_out);
    run_stream("stream cta2 A16+B16, relaxed remote arrive", true, 6, 16384, 16384, 4, n, src, src_bytes, d_out);
    run_stream("stream cta2 A16+B16, relay lane per slot", true, 6, 16384, 16384, 8, n, src, src_bytes, d_out);
    run_stream("stream cta2 A16+B16, per-slot lanes + relaxed", true, 6, 16384, 16384, 12, n, src, src_bytes, d_out);
    run_stream("stream cta2 B16 only, per-slot lanes + relaxed", true, 8, 0, 16384, 12, n, src, src_bytes, d_out);
  }
  return 0;
}
