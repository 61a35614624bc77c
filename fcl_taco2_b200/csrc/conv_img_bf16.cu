// Image-to-image k-tap convolutions / linears of the front end on tcgen05 (cta_group::2) fed by tensor-map TMA.
// Reference ops: the encoder conv stack (nets/modules/encoder_sa.py:134-140: Conv1d k5 no bias + BatchNorm eval
// (folded) + ReLU), the BiLSTM input projection (encoder_sa.py:143-146, x W_ih^T + b), the predictor stacks
// (variance_predictor.py:48-66,86-90 and the espnet DurationPredictor: Conv1d k3 + bias -> ReLU -> LayerNorm(C, eps
// 1e-12) [-> Linear(C, 1)]).
//
// Activations travel between layers as bf16 UMMA operand images in a PADDED row space (include/fcl_taco2.h):
// image[c/8][prow + 4][c%8], `gap` zero rows between utterances (every producer writes them), so
//   * a tile is 128 consecutive padded rows whatever utterances it touches (dense tiles: ~657 instead of 1046 for
//     S batch 1024), and the zero halo of each utterance is simply there;
//   * the input window of a tile for one 64-channel K stage -- stored rows [128 t, 128 t + 136) of 8 slabs -- is ONE
//     cp.async.bulk.tensor (3-D box {64 elements = 8 rows x 8 channels, 17 row groups, 8 slabs}), landing in shared
//     memory exactly as the K-major no-swizzle operand image (LBO = 2176 B, SBO = 128 B); tap t of a conv with halo h
//     is the descriptor start shifted by (4 - h + t) rows of 16 bytes.
// A CTA pair (2-CTA cluster on one TPC) owns two tiles: every weight stage is split along N between the two SMs and
// consumed by one M=256 tcgen05.mma.cta_group::2 stream issued by the leader; both CTAs' TMA loads complete on the
// LEADER's mbarrier (.cta_group::2 form), so there is no relay thread and no remote "data landed" arrival at all.
// The grid is persistent: pairs walk super-tiles; accumulators are double-buffered in TMEM when they are <= 256
// columns wide, so the epilogue of one tile overlaps the MMAs of the next.
// Round 2, second half: layers whose kchunks x taps weight stages all fit beside the window ring (the postnet: 80 KB per
// CTA) keep their weights RESIDENT in shared memory (b_resident: loaded once per CTA, no weight barriers in the issue
// loop, window ring three tiles deep), and the MMA issue loops build their descriptors with one add per MMA:
// tools/mma_floor.cu shows the pipe needs 64 cycles per M = 256, N = 128 MMA (52 for N = 64, 129 for N = 256) -- the
// ~128 cycles per MMA "whatever N" seen earlier was the issuing thread, not the pipe.
// Warp roles (640 threads): warp 0 = TMA producer, warp 1 = MMA issuer (leader), warp 2 = TMEM allocator,
// warps 4-19 = epilogue (TMEM lane quarter = warp % 4 -> row; column quarter = (warp - 4) / 4).
#include "common.cuh"
#include "umma.cuh"
#include <cuda.h>
#include <cuda_fp16.h>

namespace fcl {
namespace ci {
using namespace umma;

constexpr int kThreads = 640;
constexpr int kEpiWarps = 16;
constexpr int kMaxStages = 8;
constexpr int kMaxResident = 16;                  // weight stages of a layer kept resident (b_resident mode)
constexpr uint32_t kSlab = 136u * 16u;            // one 8-channel slab of a window: 136 rows x 16 B
constexpr uint32_t kABytes = 8u * kSlab;          // one A stage: 64 channels
constexpr int kRowOff = 4;                        // stored row = padded row + 4

struct Shared {
  uint64_t a_full[kMaxStages], a_empty[kMaxStages];
  uint64_t b_full[kMaxResident], b_empty[kMaxStages];
  uint64_t tmem_full[2], tmem_empty[2];
  uint32_t tmem_base;
  alignas(16) float ln_part[2][4][128][2];        // [tile parity][column quarter][row][sum, sum of squares]
  float head_part[2][4][128];
  alignas(16) float bias[2048];                             // per-channel epilogue parameters, staged once (broadcast LDS in the loops)
  float gamma[512], beta[512], head_w[512];
};

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a shared-memory object of THIS CTA's layout) in CTA `cta` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(cta));
  return r;
}
// "accumulator drained" from the peer's epilogue warps to the leader's barrier. No memory is published by this
// arrival (the ordering that matters is tcgen05.ld -> tcgen05.mma, carried by tcgen05.wait::ld + the tcgen05 fences),
// hence .relaxed at cluster scope.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma2_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// one arrival on the mbarrier at this offset in BOTH CTAs when all MMAs issued so far by this thread have completed
__device__ __forceinline__ void mma2_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// Tensor-map TMA loads whose completion (bytes) is signalled on the mbarrier of the LEADER CTA of the pair.
__device__ __forceinline__ void tma_load_3d_2sm(void* smem_dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

// optional timeline of CTA 0 (debug / profiling aid; p.trace == nullptr in production): {event id, clock64}.
// ids: 100 + blk (MMA issuer: accumulator slot free), 200 + blk (all MMAs of the block issued), 300 + rg (epilogue:
// slot ready), 310 + rg (pass 1 done), 320 (moments exchanged), 330 + rg (pass 2 / plain epilogue done)
__device__ __forceinline__ void ci_trace(const FclConvImgParams& p, int id) {
  if (p.trace && blockIdx.x == 0) {
    const unsigned long long n = atomicAdd(reinterpret_cast<unsigned long long*>(p.trace), 1ull);
    if (n < (unsigned long long)p.trace_cap) { p.trace[2 + 2 * n] = id; p.trace[3 + 2 * n] = clock64(); }
  }
}

template <int ACT>
__device__ __forceinline__ float act_t(float x) {
  if constexpr (ACT == FCL_ACT_RELU) return fmaxf(x, 0.f);
  else if constexpr (ACT == FCL_ACT_TANH) return tanh_fast(x);
  else return x;
}

template <int EPI, int ACT>
__global__ void __launch_bounds__(kThreads, 1)
conv_img_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap wmap, FclConvImgParams p,
                int a_stages, int b_stages, int n_pairs, int b_resident) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ Shared sh;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) ci_trace(p, 1);                         // kernel entry
  // per-CTA life span (%globaltimer, ns) behind the event records: records [trace_cap + blockIdx] = {entry, exit}
  if (p.trace && tid == 0) { unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); p.trace[2 + 2 * ((long long)p.trace_cap + blockIdx.x)] = (long long)g; }
  const uint32_t rank = cluster_ctarank();
  const int pair = (int)blockIdx.x >> 1;
  const int taps = p.taps, halo = taps >> 1, nb = p.nb;
  const int kchunks = p.cin / 64, nblk = p.cout / nb;
  constexpr bool whole_row = EPI == FCL_EPI_LN_IMAGE || EPI == FCL_EPI_LN_HEAD;      // one accumulator = the whole output row
  const int acc_cols = whole_row ? p.cout : nb;
  const int n_bufs = acc_cols <= 256 ? 2 : 1;
  // whole rows wider than 256 columns (predictors: 384 = 2 N blocks): the two N blocks are separate accumulator SLOTS
  // with their own full/empty barriers, so the epilogue's first pass over block 0 overlaps the MMAs of block 1 and the
  // next tile's block 0 starts as soon as block 0 has been normalised.
  const bool split = whole_row && n_bufs == 1;
  const uint32_t b_bytes = (uint32_t)(nb / 2) * 128u;        // this CTA's half of a weight stage: nb/2 columns x 64 k
  const int n_super = (p.n_tiles + 1) >> 1;
  const long rows_alloc = (long)p.n_tiles * 128 + 8;
  uint8_t* a_ring = smem;
  uint8_t* b_ring = smem + (size_t)a_stages * kABytes;

  if (tid == 0) {
    for (int s = 0; s < a_stages; ++s) { mbar_init(&sh.a_full[s], 1); mbar_init(&sh.a_empty[s], 1); }
    for (int s = 0; s < b_stages; ++s) { mbar_init(&sh.b_full[s], 1); if (s < kMaxStages) mbar_init(&sh.b_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&sh.tmem_full[b], 1); mbar_init(&sh.tmem_empty[b], 2 * kEpiWarps); }
    fence_barrier_init();
    prefetch_tmap(&amap);
    prefetch_tmap(&wmap);
  }
  for (int i = tid; i < p.cout; i += kThreads) {
    sh.bias[i] = p.bias ? __ldg(p.bias + i) : 0.f;
    if (whole_row) {
      sh.gamma[i] = __ldg(p.gamma + i);
      sh.beta[i] = __ldg(p.beta + i);
      sh.head_w[i] = p.head_w ? __ldg(p.head_w + i) : 0.f;
    }
  }
  cluster_sync_all();                                   // barriers of both CTAs initialised before any remote signal
  if (warp == 2) tmem_alloc2(&sh.tmem_base, 512);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = sh.tmem_base;
  if (tid == 0) ci_trace(p, 2);                         // set-up done

  if (warp == 0) {
    // ================================================================ TMA producer (both CTAs)
    if (b_resident && elect_one()) {
      // RESIDENT WEIGHTS (one N block, all kchunks * taps weight stages of the layer fit beside the window ring: the
      // postnet layers, 80 KB per CTA): loaded once, every tile re-reads them in place. No weight traffic after the
      // first tile, no weight barriers in the MMA issuer's loop, and the window ring gets the remaining shared memory
      // (three tiles of windows in flight).
      if (pair < n_super) {
        for (int i = 0; i < kchunks * taps; ++i) {
          if (rank == 0) mbar_arrive_expect_tx(&sh.b_full[i], 2u * b_bytes);
          tma_load_2d_2sm(b_ring + (size_t)i * b_bytes, &wmap, mapa_u32(&sh.b_full[i], 0), 0, (i * 2 + (int)rank) * (nb / 4));
        }
      }
      uint32_t a_ctr = 0;
      for (int st = pair; st < n_super; st += n_pairs) {
        const int tile = 2 * st + (int)rank;
        for (int kc = 0; kc < kchunks; ++kc) {
          const uint32_t s = a_ctr % (uint32_t)a_stages, ph = (a_ctr / (uint32_t)a_stages) & 1u;
          mbar_wait(&sh.a_empty[s], ph ^ 1u);
          if (rank == 0) mbar_arrive_expect_tx(&sh.a_full[s], 2u * kABytes);
          tma_load_3d_2sm(a_ring + (size_t)s * kABytes, &amap, mapa_u32(&sh.a_full[s], 0), 0, tile * 16, kc * 8);
          ++a_ctr;
        }
      }
    } else if (!b_resident && elect_one()) {
      uint32_t a_ctr = 0, b_ctr = 0;
      for (int st = pair; st < n_super; st += n_pairs) {
        const int tile = 2 * st + (int)rank;            // an odd tile count leaves the last peer a tile past the end: TMA zero-fills
        for (int blk = 0; blk < nblk; ++blk) {
          for (int kc = 0; kc < kchunks; ++kc) {
            if (blk == 0) {                             // the window is loaded once per tile and re-read by later N blocks
              const uint32_t s = a_ctr % (uint32_t)a_stages, ph = (a_ctr / (uint32_t)a_stages) & 1u;
              mbar_wait(&sh.a_empty[s], ph ^ 1u);
              if (rank == 0) mbar_arrive_expect_tx(&sh.a_full[s], 2u * kABytes);
              tma_load_3d_2sm(a_ring + (size_t)s * kABytes, &amap, mapa_u32(&sh.a_full[s], 0), 0, tile * 16, kc * 8);
              ++a_ctr;
            }
            for (int t = 0; t < taps; ++t) {
              const uint32_t s = b_ctr % (uint32_t)b_stages, ph = (b_ctr / (uint32_t)b_stages) & 1u;
              mbar_wait(&sh.b_empty[s], ph ^ 1u);
              if (rank == 0) mbar_arrive_expect_tx(&sh.b_full[s], 2u * b_bytes);
              const int block = ((blk * kchunks + kc) * taps + t) * 2 + (int)rank;       // [blk][kc][tap][half]
              tma_load_2d_2sm(b_ring + (size_t)s * b_bytes, &wmap, mapa_u32(&sh.b_full[s], 0), 0, block * (nb / 4));
              ++b_ctr;
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================================================ MMA issuer (leader CTA only)
    if (b_resident && rank == 0 && elect_one()) {
      // resident weights: nblk == 1, accumulators double-buffered (launcher guarantees acc_cols <= 256).
      // This thread's instruction chain is what bounds the layer: tools/mma_floor.cu measures 64 cycles per dependent
      // M = 256, N = 128 MMA from a lean issue loop (52 for N = 64, 129 for N = 256) -- there is no 128-cycle floor in
      // the pipe, the loops of round 2 simply took >= 120 cycles per MMA to issue. Hence: descriptors as
      // (constant high word, low word = one add), the first tile (weights still landing) peeled off, everything that
      // is loop-invariant hoisted.
      uint32_t a_ctr = 0, acc_ctr = 0;
      const uint32_t idesc = idesc_op_f32(256u, (uint32_t)nb);
      const uint32_t b_fld = ((uint32_t)(nb / 2)) << 16, b_kstep = (uint32_t)nb;       // LBO field, K = 16 step (>> 4)
      const uint32_t b_lo0 = (smem_u32(b_ring) >> 4) | b_fld, b_stage16 = b_bytes >> 4;
      const uint32_t a_ring_lo = (smem_u32(a_ring) >> 4) + (uint32_t)(kRowOff - halo), a_fld = (kSlab >> 4) << 16;
      constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);
      constexpr uint32_t kAStep = (2u * kSlab) >> 4;
      for (int st = pair; st < n_super; st += n_pairs) {
        const uint32_t buf = acc_ctr & 1u, use = acc_ctr >> 1;
        mbar_wait(&sh.tmem_empty[buf], (use & 1u) ^ 1u);
        tc_fence_after();
        ci_trace(p, 100);
        const uint32_t d_tmem = tmem + buf * 256u;
        uint32_t acc = 0u, b_lo = b_lo0;
        for (int kc = 0; kc < kchunks; ++kc) {
          const uint32_t sa = a_ctr % (uint32_t)a_stages;
          mbar_wait(&sh.a_full[sa], (a_ctr / (uint32_t)a_stages) & 1u);
          tc_fence_after();
          uint32_t a_lo = (a_ring_lo + sa * (kABytes >> 4)) | a_fld;
          if (st == pair) {                              // first tile of this pair: the weights are still landing
            for (int t = 0; t < taps; ++t) {
              mbar_wait(&sh.b_full[kc * taps + t], 0u);
              tc_fence_after();
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                mma2_bf16_ss(d_tmem, ((uint64_t)kDescHi << 32) | (a_lo + (uint32_t)k * kAStep),
                             ((uint64_t)kDescHi << 32) | (b_lo + (uint32_t)k * b_kstep), idesc, acc);
                acc = 1u;
              }
              a_lo += 1u;                                // next tap: one row (16 bytes) further
              b_lo += b_stage16;
            }
          } else {
#pragma unroll 1
            for (int t = 0; t < taps; ++t) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                mma2_bf16_ss(d_tmem, ((uint64_t)kDescHi << 32) | (a_lo + (uint32_t)k * kAStep),
                             ((uint64_t)kDescHi << 32) | (b_lo + (uint32_t)k * b_kstep), idesc, acc);
                acc = 1u;
              }
              a_lo += 1u;
              b_lo += b_stage16;
            }
          }
          mma2_commit(&sh.a_empty[sa]);
          ++a_ctr;
        }
        mma2_commit(&sh.tmem_full[buf]);
        ci_trace(p, 200);
        ++acc_ctr;
      }
    } else if (!b_resident && rank == 0 && elect_one()) {
      uint32_t a_ctr = 0, acc_ctr = 0;
      uint32_t sb = 0, sb_par = 0;                       // weight ring position / parity
      const uint32_t idesc = idesc_op_f32(256u, (uint32_t)nb);
      const uint32_t b_fld = ((uint32_t)(nb / 2)) << 16, b_kstep = (uint32_t)nb;       // LBO field, K = 16 step (>> 4)
      const uint32_t b_ring_lo = smem_u32(b_ring) >> 4, b_stage16 = b_bytes >> 4;
      const uint32_t a_ring_lo = (smem_u32(a_ring) >> 4) + (uint32_t)(kRowOff - halo), a_fld = (kSlab >> 4) << 16;
      constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);
      constexpr uint32_t kAStep = (2u * kSlab) >> 4;
      for (int st = pair; st < n_super; st += n_pairs) {
        const uint32_t a_base = a_ctr;
        for (int blk = 0; blk < nblk; ++blk) {
          const uint32_t buf = split ? (uint32_t)blk : acc_ctr % (uint32_t)n_bufs;
          const uint32_t use = split ? acc_ctr : acc_ctr / (uint32_t)n_bufs;
          if (blk == 0 || !whole_row || split) {
            mbar_wait(&sh.tmem_empty[buf], (use & 1u) ^ 1u);
            tc_fence_after();
            ci_trace(p, 100 + blk);
          }
          const uint32_t d_tmem = split ? tmem + (uint32_t)(blk * nb) : tmem + buf * 256u + (whole_row ? (uint32_t)(blk * nb) : 0u);
          for (int kc = 0; kc < kchunks; ++kc) {
            const uint32_t sa = (a_base + (uint32_t)kc) % (uint32_t)a_stages;
            if (blk == 0) {
              mbar_wait(&sh.a_full[sa], ((a_base + (uint32_t)kc) / (uint32_t)a_stages) & 1u);
              tc_fence_after();
            }
            // descriptors: constant high word, low word = (address >> 4) | (LBO >> 4) << 16 advanced by adds (see the
            // resident loop above: this thread's instruction chain per MMA is what bounds the N < 256 layers)
            uint32_t a_lo = ((a_ring_lo + sa * (kABytes >> 4)) | a_fld);
            uint32_t acc = kc > 0 ? 1u : 0u;
            for (int t = 0; t < taps; ++t) {
              mbar_wait(&sh.b_full[sb], sb_par);
              tc_fence_after();
              const uint32_t b_lo = (b_ring_lo + sb * b_stage16) | b_fld;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                mma2_bf16_ss(d_tmem, ((uint64_t)kDescHi << 32) | (a_lo + (uint32_t)k * kAStep),
                             ((uint64_t)kDescHi << 32) | (b_lo + (uint32_t)k * b_kstep), idesc, acc);
                acc = 1u;
              }
              a_lo += 1u;                               // next tap: one row (16 bytes) further
              mma2_commit(&sh.b_empty[sb]);             // frees the weight stage in BOTH CTAs
              if (++sb == (uint32_t)b_stages) { sb = 0; sb_par ^= 1u; }
            }
            if (blk == nblk - 1) mma2_commit(&sh.a_empty[sa]);     // the window stage is free once the last N block used it
          }
          if (blk == nblk - 1 || !whole_row || split) {
            mma2_commit(&sh.tmem_full[buf]);            // accumulator (slot) ready in BOTH CTAs
            ci_trace(p, 200 + blk);
            if (!split || blk == nblk - 1) ++acc_ctr;
          }
        }
        a_ctr += (uint32_t)kchunks;
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ================================================================ epilogue (512 threads of each CTA, own 128 rows)
    // Instruction-lean on purpose: with 16 warps the SM's four schedulers issue ~4 k warp-instructions per tile here,
    // and in the LayerNorm variants part of that sits on the tile's critical path (the epilogue kind and the
    // activation are template parameters, parameters come from shared memory as float4, addresses are incremental).
    const int q = warp & 3, cs = (warp - 4) >> 2;
    const int r = q * 32 + lane;                        // row within the tile == TMEM lane
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
    const int cq = acc_cols / 4;                        // accumulator columns per thread (multiple of 16)
    const uint32_t empty_leader = mapa_u32(&sh.tmem_empty[0], 0);
    uint32_t acc_ctr = 0, tile_ctr = 0;
    uint8_t* oimg = reinterpret_cast<uint8_t*>(p.out_img);
    const size_t slab_stride = (size_t)rows_alloc * 16;
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    auto release = [&](uint32_t slot) {                 // this warp has drained accumulator slot `slot`
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(&sh.tmem_empty[slot]);
        else mbar_arrive_cluster(empty_leader + slot * (uint32_t)sizeof(uint64_t));
      }
    };
    for (int st = pair; st < n_super; st += n_pairs, ++tile_ctr) {
      const int tile = 2 * st + (int)rank;
      const bool tile_ok = tile < p.n_tiles;
      const long prow = (long)tile * 128 + r;
      const int src = tile_ok ? __ldg(p.prow_src + prow) : -1;          // original row, -1 = gap row
      const bool live = src >= 0;
      uint8_t* orow = oimg + (size_t)(prow + kRowOff) * 16;             // this row's 16-byte line in slab 0
      // guard rows of the output image that no tile owns: stored rows [0, 4) and [128 T + 4, 128 T + 8)
      long zrow = -1;
      if (tile == 0 && r < kRowOff) zrow = r;
      else if (tile == p.n_tiles - 1 && r >= 128 - kRowOff) zrow = (long)p.n_tiles * 128 + kRowOff + (r - (128 - kRowOff));
      if constexpr (EPI == FCL_EPI_IMAGE || EPI == FCL_EPI_BLOCKED_F32 || EPI == FCL_EPI_BLOCKED_F16 || EPI == FCL_EPI_ROWS_F32) {
        const int n_acc = whole_row ? 1 : nblk;
        for (int a = 0; a < n_acc; ++a) {
          const uint32_t buf = acc_ctr % (uint32_t)n_bufs, use = acc_ctr / (uint32_t)n_bufs;
          const int c_base = a * nb + cs * cq;                          // first output channel of this thread
          const uint32_t t_addr = lane_addr + buf * 256u + (uint32_t)(cs * cq);
          mbar_wait(&sh.tmem_full[buf], use & 1u);
          tc_fence_after();
          if (tid == 128) ci_trace(p, 300 + a);
#pragma unroll 1
          for (int g = 0; g < cq / 16; ++g) {
            float v[16];
            tmem_ld16(t_addr + (uint32_t)(g * 16), v);
            const int c0 = c_base + g * 16;
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
              const float4 b = *reinterpret_cast<const float4*>(sh.bias + c0 + 4 * qd);
              v[4 * qd] = act_t<ACT>(v[4 * qd] + b.x); v[4 * qd + 1] = act_t<ACT>(v[4 * qd + 1] + b.y);
              v[4 * qd + 2] = act_t<ACT>(v[4 * qd + 2] + b.z); v[4 * qd + 3] = act_t<ACT>(v[4 * qd + 3] + b.w);
            }
            if constexpr (EPI == FCL_EPI_IMAGE) {
              uint4 w0 = make_uint4(pack_op(v[0], v[1]), pack_op(v[2], v[3]), pack_op(v[4], v[5]), pack_op(v[6], v[7]));
              uint4 w1 = make_uint4(pack_op(v[8], v[9]), pack_op(v[10], v[11]), pack_op(v[12], v[13]), pack_op(v[14], v[15]));
              if (!live) { w0 = zero4; w1 = zero4; }
              if (tile_ok) {
                uint8_t* o = orow + (size_t)(c0 >> 3) * slab_stride;
                *reinterpret_cast<uint4*>(o) = w0;
                *reinterpret_cast<uint4*>(o + slab_stride) = w1;
              }
            } else if constexpr (EPI == FCL_EPI_ROWS_F32) {
              if (live && c0 < p.out_chans) {
                float4* o = reinterpret_cast<float4*>(p.out_rows + (size_t)src * p.ldo + c0);
                const float4* res = p.residual ? reinterpret_cast<const float4*>(p.residual + (size_t)src * p.ldr + c0) : nullptr;
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) {
                  if (c0 + 4 * qd < p.out_chans) {
                    float4 x = make_float4(v[4 * qd], v[4 * qd + 1], v[4 * qd + 2], v[4 * qd + 3]);
                    if (res) { const float4 rr = __ldg(res + qd); x.x += rr.x; x.y += rr.y; x.z += rr.z; x.w += rr.w; }
                    o[qd] = x;
                  }
                }
              }
            } else if constexpr (EPI == FCL_EPI_BLOCKED_F32) {
              if (live) {
                float4* o = reinterpret_cast<float4*>(p.out_blk + ((size_t)(c0 >> 4) * (size_t)p.n_tiles * 128 + (size_t)prow) * 16);
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) o[qd] = make_float4(v[4 * qd], v[4 * qd + 1], v[4 * qd + 2], v[4 * qd + 3]);
              }
            } else {
              if (live) {
                uint32_t h[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                  const __half2 hh = __floats2half2_rn(fminf(fmaxf(v[2 * k], -65504.f), 65504.f), fminf(fmaxf(v[2 * k + 1], -65504.f), 65504.f));
                  h[k] = *reinterpret_cast<const uint32_t*>(&hh);
                }
                uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.out_blk) + ((size_t)(c0 >> 4) * (size_t)p.n_tiles * 128 + (size_t)prow) * 16);
                o[0] = make_uint4(h[0], h[1], h[2], h[3]);
                o[1] = make_uint4(h[4], h[5], h[6], h[7]);
              }
            }
          }
          release(buf);
          if (tid == 128) ci_trace(p, 330 + a);
          ++acc_ctr;
          if (EPI == FCL_EPI_IMAGE && zrow >= 0)                        // rare: first / last tile only
            for (int k8 = 0; k8 < cq / 8; ++k8) *reinterpret_cast<uint4*>(oimg + (size_t)((c_base >> 3) + k8) * slab_stride + (size_t)zrow * 16) = zero4;
        }
      } else {
        // ---- LayerNorm over the whole row: 4 threads per row, each owns `cr` columns of every range (one range, or one
        // per N block in split mode). Pass 1 = moments; pass 2 = normalise, reading TMEM again (cheaper than holding the
        // row in registers). In split mode pass 1 of block 0 runs under the MMAs of block 1, and block 0 is handed back
        // to the MMA issuer before block 1 is normalised.
        const uint32_t buf = split ? 0u : acc_ctr % (uint32_t)n_bufs;
        const uint32_t use = split ? acc_ctr : acc_ctr / (uint32_t)n_bufs;
        const uint32_t t_base = lane_addr + (split ? 0u : buf * 256u);
        const int par = (int)(tile_ctr & 1u);
        const int n_rng = split ? nblk : 1;
        const int cr = split ? nb / 4 : cq;
        float s1 = 0.f, s2 = 0.f;
        for (int rg = 0; rg < n_rng; ++rg) {
          mbar_wait(&sh.tmem_full[split ? (uint32_t)rg : buf], use & 1u);
          tc_fence_after();
          if (tid == 128) ci_trace(p, 300 + rg);
          const int cb = (split ? rg * nb : 0) + cs * cr;
#pragma unroll 1
          for (int g = 0; g < cr / 16; ++g) {
            float v[16];
            tmem_ld16(t_base + (uint32_t)(cb + g * 16), v);
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
              const float4 b = *reinterpret_cast<const float4*>(sh.bias + cb + g * 16 + 4 * qd);
              const float x0 = act_t<ACT>(v[4 * qd] + b.x), x1 = act_t<ACT>(v[4 * qd + 1] + b.y);
              const float x2 = act_t<ACT>(v[4 * qd + 2] + b.z), x3 = act_t<ACT>(v[4 * qd + 3] + b.w);
              s1 += (x0 + x1) + (x2 + x3);
              s2 = fmaf(x0, x0, fmaf(x1, x1, fmaf(x2, x2, fmaf(x3, x3, s2))));
            }
          }
          if (tid == 128) ci_trace(p, 310 + rg);
        }
        sh.ln_part[par][cs][r][0] = s1;
        sh.ln_part[par][cs][r][1] = s2;
        epi_bar_sync();
        if (tid == 128) ci_trace(p, 320);
        float t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) { t1 += sh.ln_part[par][k][r][0]; t2 += sh.ln_part[par][k][r][1]; }
        const float inv_c = 1.0f / (float)p.cout;
        const float mean = t1 * inv_c;
        const float var = fmaxf(t2 * inv_c - mean * mean, 0.f);          // biased variance (torch layer_norm)
        const float rstd = 1.0f / sqrtf(var + 1e-12f);
        const float nmr = -mean * rstd;                                  // (x - mean) * rstd = fma(x, rstd, nmr)
        float dot = 0.f;
        for (int rg = 0; rg < n_rng; ++rg) {
          const int cb = (split ? rg * nb : 0) + cs * cr;
#pragma unroll 1
          for (int g = 0; g < cr / 16; ++g) {
            float v[16];
            tmem_ld16(t_base + (uint32_t)(cb + g * 16), v);
            const int c0 = cb + g * 16;
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
              const float4 b = *reinterpret_cast<const float4*>(sh.bias + c0 + 4 * qd);
              const float4 ga = *reinterpret_cast<const float4*>(sh.gamma + c0 + 4 * qd);
              const float4 be = *reinterpret_cast<const float4*>(sh.beta + c0 + 4 * qd);
              v[4 * qd] = fmaf(fmaf(act_t<ACT>(v[4 * qd] + b.x), rstd, nmr), ga.x, be.x);
              v[4 * qd + 1] = fmaf(fmaf(act_t<ACT>(v[4 * qd + 1] + b.y), rstd, nmr), ga.y, be.y);
              v[4 * qd + 2] = fmaf(fmaf(act_t<ACT>(v[4 * qd + 2] + b.z), rstd, nmr), ga.z, be.z);
              v[4 * qd + 3] = fmaf(fmaf(act_t<ACT>(v[4 * qd + 3] + b.w), rstd, nmr), ga.w, be.w);
            }
            if constexpr (EPI == FCL_EPI_LN_HEAD) {
#pragma unroll
              for (int qd = 0; qd < 4; ++qd) {
                const float4 hw = *reinterpret_cast<const float4*>(sh.head_w + c0 + 4 * qd);
                dot = fmaf(v[4 * qd], hw.x, fmaf(v[4 * qd + 1], hw.y, fmaf(v[4 * qd + 2], hw.z, fmaf(v[4 * qd + 3], hw.w, dot))));
              }
            } else {
              uint4 w0 = make_uint4(pack_op(v[0], v[1]), pack_op(v[2], v[3]), pack_op(v[4], v[5]), pack_op(v[6], v[7]));
              uint4 w1 = make_uint4(pack_op(v[8], v[9]), pack_op(v[10], v[11]), pack_op(v[12], v[13]), pack_op(v[14], v[15]));
              if (!live) { w0 = zero4; w1 = zero4; }
              if (tile_ok) {
                uint8_t* o = orow + (size_t)(c0 >> 3) * slab_stride;
                *reinterpret_cast<uint4*>(o) = w0;
                *reinterpret_cast<uint4*>(o + slab_stride) = w1;
              }
            }
          }
          release(split ? (uint32_t)rg : buf);             // block `rg` is drained: the next tile's MMAs may overwrite it
          if (tid == 128) ci_trace(p, 330 + rg);
          if (EPI == FCL_EPI_LN_IMAGE && zrow >= 0)
            for (int k8 = 0; k8 < cr / 8; ++k8) *reinterpret_cast<uint4*>(oimg + (size_t)((cb >> 3) + k8) * slab_stride + (size_t)zrow * 16) = zero4;
        }
        ++acc_ctr;
        if constexpr (EPI == FCL_EPI_LN_HEAD) {
          sh.head_part[par][cs][r] = dot;
          epi_bar_sync();
          if (cs == 0 && live) {
            const float h = ((sh.head_part[par][0][r] + sh.head_part[par][1][r]) + (sh.head_part[par][2][r] + sh.head_part[par][3][r])) + p.head_b;
            if (p.head_out) p.head_out[src] = h;
            if (p.dur_out) {
              // clamp(round_half_even(exp(x) - 1), 0, cap); rintf rounds half to even like torch.round
              float d = rintf(expf(h) - 1.0f);
              d = fminf(fmaxf(d, 0.f), (float)FCL_MAX_DURATION);
              p.dur_out[src] = (int)d;
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  if (tid == 0) ci_trace(p, 3);                         // producer done
  cluster_sync_all();                                   // the leader's MMAs read the peer's shared memory: leave together
  if (warp == 2) tmem_dealloc2(tmem, 512);
  if (tid == 0) ci_trace(p, 4);                         // exit
  if (p.trace && tid == 0) { unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); p.trace[3 + 2 * ((long long)p.trace_cap + blockIdx.x)] = (long long)g; }
}

// ---------------------------------------------------------------- padded row space
__global__ void __launch_bounds__(256)
pad_rows_kernel(FclPadRowsParams p) {
  const int total = p.n_tiles * 128;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total + p.n_utts + 1; i += gridDim.x * blockDim.x) {
    if (i >= total) {                                   // padded offset of every utterance
      const int u = i - total;
      p.prow_off[u] = p.utt_off[u] + p.gap * u;
      continue;
    }
    int lo = 0, hi = p.n_utts;                          // last utterance whose padded offset <= i
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(p.utt_off + mid) + p.gap * mid <= i) lo = mid; else hi = mid;
    }
    const int o = __ldg(p.utt_off + lo), len = __ldg(p.utt_off + lo + 1) - o;
    const int j = i - (o + p.gap * lo);
    p.prow_src[i] = (j >= 0 && j < len) ? o + j : -1;
  }
}

// fp32 rows (optionally gathered through embedding ids) -> bf16 image in the padded row space, zero elsewhere
__global__ void __launch_bounds__(256)
rows_to_image_kernel(FclRowsToImageParams p) {
  const long rows_alloc = (long)p.n_tiles * 128 + 8;
  const int slabs = p.chans >> 3;
  const int src_chans = p.src_chans > 0 ? p.src_chans : p.chans;
  const long total = rows_alloc * slabs;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int slab = (int)(i / rows_alloc);
    const long sr = i - (long)slab * rows_alloc;
    const long prow = sr - kRowOff;
    int src = -1;
    if (prow >= 0 && prow < (long)p.n_tiles * 128) src = __ldg(p.prow_src + prow);
    uint4 w = make_uint4(0u, 0u, 0u, 0u);
    if (src >= 0 && slab * 8 < src_chans) {
      const size_t row = p.gather ? (size_t)__ldg(p.gather + src) : (size_t)src;
      const float4* s = reinterpret_cast<const float4*>(p.src + row * p.ld + (size_t)slab * 8);
      const float4 a = __ldg(s), b = __ldg(s + 1);
      w = make_uint4(umma::pack_op(a.x, a.y), umma::pack_op(a.z, a.w), umma::pack_op(b.x, b.y), umma::pack_op(b.z, b.w));
    }
    reinterpret_cast<uint4*>(p.img)[i] = w;
  }
}

#ifdef FCL_OPERANDS_BF16
constexpr CUtensorMapDataType kTmapDtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
#else
constexpr CUtensorMapDataType kTmapDtype = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
#endif
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

}  // namespace ci
}  // namespace fcl

extern "C" int fcl_pad_rows(const FclPadRowsParams* p, void* stream) {
  using namespace fcl;
  FCL_REQUIRE(p && p->utt_off && p->prow_src && p->prow_off, "null pointer");
  FCL_REQUIRE(p->n_utts > 0 && p->n_rows > 0 && p->gap >= 0 && p->gap <= 4, "bad sizes");
  FCL_REQUIRE((long)p->n_tiles * 128 >= (long)p->n_rows + (long)p->gap * (p->n_utts - 1), "n_tiles too small for the padded row space");
  const int total = p->n_tiles * 128 + p->n_utts + 1;
  ci::pad_rows_kernel<<<(total + 255) / 256, 256, 0, as_stream(stream)>>>(*p);
  return check_launch("fcl_pad_rows");
}

extern "C" int fcl_rows_to_image(const FclRowsToImageParams* p, void* stream) {
  using namespace fcl;
  FCL_REQUIRE(p && p->src && p->prow_src && p->img, "null pointer");
  FCL_REQUIRE(p->n_tiles > 0 && p->chans > 0 && p->chans % 8 == 0 && p->ld % 4 == 0, "bad sizes");
  FCL_REQUIRE(p->src_chans >= 0 && p->src_chans <= p->chans && p->src_chans % 8 == 0, "src_chans must be a multiple of 8, <= chans");
  const long total = ((long)p->n_tiles * 128 + 8) * (p->chans / 8);
  int sms = fcl_sm_count();
  if (sms < 0) return sms;
  const long blocks = (total + 255) / 256;
  ci::rows_to_image_kernel<<<(unsigned)(blocks < (long)sms * 16 ? blocks : (long)sms * 16), 256, 0, as_stream(stream)>>>(*p);
  return check_launch("fcl_rows_to_image");
}

extern "C" int fcl_conv_img_bf16(const FclConvImgParams* p, void* stream) {
  using namespace fcl;
  using namespace fcl::ci;
  FCL_REQUIRE(p && p->in_img && p->w_packed && p->prow_src, "null pointer");
  FCL_REQUIRE(p->n_tiles > 0 && p->cin >= 64 && p->cin % 64 == 0 && p->cin <= 512, "cin must be a multiple of 64, <= 512");
  FCL_REQUIRE(p->taps == 1 || p->taps == 3 || p->taps == 5, "taps must be 1, 3 or 5");
  FCL_REQUIRE(p->nb >= 64 && p->nb <= 256 && p->nb % 64 == 0 && p->cout % p->nb == 0, "nb must be a multiple of 64 (<= 256) dividing cout");
  FCL_REQUIRE(p->cout <= 2048, "cout must be <= 2048");
  const bool ln = p->epi == FCL_EPI_LN_IMAGE || p->epi == FCL_EPI_LN_HEAD;
  FCL_REQUIRE(p->epi >= FCL_EPI_IMAGE && p->epi <= FCL_EPI_ROWS_F32, "unknown epilogue");
  FCL_REQUIRE(p->epi != FCL_EPI_ROWS_F32 || (p->out_rows && p->out_chans > 0 && p->out_chans <= p->cout && p->out_chans % 4 == 0 &&
                                              p->ldo % 4 == 0 && (!p->residual || p->ldr % 4 == 0)),
              "row epilogue needs out_rows, out_chans (multiple of 4, <= cout) and leading dimensions that are multiples of 4");
  FCL_REQUIRE(!ln || (p->cout <= 512 && p->gamma && p->beta), "LayerNorm epilogues need gamma/beta and cout <= 512");
  FCL_REQUIRE(!ln || p->cout <= 256 || p->cout == 2 * p->nb, "LayerNorm rows wider than 256 columns must be exactly two N blocks");
  FCL_REQUIRE(p->epi != FCL_EPI_LN_HEAD || (p->head_w && (p->head_out || p->dur_out)), "head epilogue needs head_w and an output");
  FCL_REQUIRE((p->epi != FCL_EPI_IMAGE && p->epi != FCL_EPI_LN_IMAGE) || p->out_img, "image epilogues need out_img");
  FCL_REQUIRE((p->epi != FCL_EPI_BLOCKED_F32 && p->epi != FCL_EPI_BLOCKED_F16) || p->out_blk, "blocked epilogues need out_blk");
  EncodeTiledFn enc = encode_fn();
  if (!enc) { set_error("fcl_conv_img_bf16: cuTensorMapEncodeTiled is not available from this driver"); return FCL_EUNSUPPORTED; }
  const int kchunks = p->cin / 64, nblk = p->cout / p->nb;
  const long rows_alloc = (long)p->n_tiles * 128 + 8;
  // input image: [cin/8 slabs][rows_alloc/8 groups][64 elements = 8 rows x 8 channels]
  CUtensorMap amap, wmap;
  {
    cuuint64_t dims[3] = {64, (cuuint64_t)(rows_alloc / 8), (cuuint64_t)(p->cin / 8)};
    cuuint64_t strides[2] = {128, (cuuint64_t)rows_alloc * 16};
    cuuint32_t box[3] = {64, 17, 8}, es[3] = {1, 1, 1};
    CUresult r = enc(&amap, kTmapDtype, 3, const_cast<void*>(p->in_img), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("fcl_conv_img_bf16: cuTensorMapEncodeTiled(input image) failed (%d)", (int)r); return FCL_ECUDA; }
  }
  {
    // weights: contiguous half-stage blocks of (nb/2) x 64 bf16 = nb/4 rows of 256 bytes
    const cuuint64_t blocks = (cuuint64_t)nblk * kchunks * p->taps * 2;
    cuuint64_t dims[2] = {128, blocks * (cuuint64_t)(p->nb / 4)};
    cuuint64_t strides[1] = {256};
    cuuint32_t box[2] = {128, (cuuint32_t)(p->nb / 4)}, es[2] = {1, 1};
    CUresult r = enc(&wmap, kTmapDtype, 2, const_cast<void*>(p->w_packed), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("fcl_conv_img_bf16: cuTensorMapEncodeTiled(weights) failed (%d)", (int)r); return FCL_ECUDA; }
  }
  // shared memory: the window of a tile stays resident while several N blocks re-read it
  const size_t b_bytes = (size_t)(p->nb / 2) * 128;
  int a_stages = nblk > 1 ? kchunks : (kchunks < 4 ? kchunks + 1 : 4);
  if (a_stages > kMaxStages) a_stages = kMaxStages;
  FCL_REQUIRE(nblk == 1 || a_stages >= kchunks, "cin too wide to keep a tile's window resident");
  const size_t budget = 192 * 1024;
  int b_stages = (int)((budget - (size_t)a_stages * kABytes) / b_bytes);
  if (b_stages > kMaxStages) b_stages = kMaxStages;
  FCL_REQUIRE(b_stages >= 2, "shared memory budget exceeded");
  // resident weights (see the kernel): one N block of at most 256 columns whose kchunks * taps stages all fit
  const int total_b = kchunks * p->taps;
  const bool ln_epi = p->epi == FCL_EPI_LN_IMAGE || p->epi == FCL_EPI_LN_HEAD;
  int b_resident = 0;
  if (nblk == 1 && !ln_epi && p->epi != FCL_EPI_ROWS_F32 && p->nb <= 256 && total_b >= 2 && total_b <= kMaxResident &&
      (size_t)total_b * b_bytes + (size_t)a_stages * kABytes <= budget) {
    b_resident = 1;
    b_stages = total_b;
    a_stages = (int)((budget - (size_t)total_b * b_bytes) / kABytes);
    if (a_stages > 3 * kchunks) a_stages = 3 * kchunks;        // three tiles of windows in flight
    if (a_stages > kMaxStages) a_stages = kMaxStages;
  }
  const size_t smem = (size_t)a_stages * kABytes + (size_t)b_stages * b_bytes + 1024;      // + alignment slack
  typedef void (*Kern)(const CUtensorMap, const CUtensorMap, FclConvImgParams, int, int, int, int);
  Kern kern = nullptr;
  const int act = p->act;
  if (p->epi == FCL_EPI_IMAGE) kern = act == FCL_ACT_RELU ? conv_img_kernel<FCL_EPI_IMAGE, FCL_ACT_RELU> : act == FCL_ACT_TANH ? conv_img_kernel<FCL_EPI_IMAGE, FCL_ACT_TANH> : conv_img_kernel<FCL_EPI_IMAGE, FCL_ACT_NONE>;
  else if (p->epi == FCL_EPI_LN_IMAGE) kern = act == FCL_ACT_RELU ? conv_img_kernel<FCL_EPI_LN_IMAGE, FCL_ACT_RELU> : conv_img_kernel<FCL_EPI_LN_IMAGE, FCL_ACT_NONE>;
  else if (p->epi == FCL_EPI_LN_HEAD) kern = act == FCL_ACT_RELU ? conv_img_kernel<FCL_EPI_LN_HEAD, FCL_ACT_RELU> : conv_img_kernel<FCL_EPI_LN_HEAD, FCL_ACT_NONE>;
  else if (p->epi == FCL_EPI_BLOCKED_F32) kern = act == FCL_ACT_RELU ? conv_img_kernel<FCL_EPI_BLOCKED_F32, FCL_ACT_RELU> : conv_img_kernel<FCL_EPI_BLOCKED_F32, FCL_ACT_NONE>;
  else if (p->epi == FCL_EPI_ROWS_F32) kern = act == FCL_ACT_RELU ? conv_img_kernel<FCL_EPI_ROWS_F32, FCL_ACT_RELU> : act == FCL_ACT_TANH ? conv_img_kernel<FCL_EPI_ROWS_F32, FCL_ACT_TANH> : conv_img_kernel<FCL_EPI_ROWS_F32, FCL_ACT_NONE>;
  else kern = act == FCL_ACT_RELU ? conv_img_kernel<FCL_EPI_BLOCKED_F16, FCL_ACT_RELU> : conv_img_kernel<FCL_EPI_BLOCKED_F16, FCL_ACT_NONE>;
  FCL_REQUIRE(act == FCL_ACT_NONE || act == FCL_ACT_RELU || (act == FCL_ACT_TANH && (p->epi == FCL_EPI_IMAGE || p->epi == FCL_EPI_ROWS_F32)), "unsupported activation for this epilogue");
  if (int rc = ensure_dyn_smem_fn(reinterpret_cast<const void*>(kern), smem, "fcl_conv_img_bf16")) return rc;
  int sms = fcl_sm_count();
  if (sms < 0) return sms;
  const int n_super = (p->n_tiles + 1) / 2;
  int n_pairs = p->n_pairs > 0 ? p->n_pairs : sms / 2;
  if (n_pairs > n_super) n_pairs = n_super;
  if (n_pairs > sms / 2) n_pairs = sms / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * n_pairs));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = as_stream(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, amap, wmap, *p, a_stages, b_stages, n_pairs, b_resident);
  if (e != cudaSuccess) { set_error("fcl_conv_img_bf16: %s", cudaGetErrorString(e)); return FCL_ECUDA; }
  return check_launch("fcl_conv_img_bf16");
}
