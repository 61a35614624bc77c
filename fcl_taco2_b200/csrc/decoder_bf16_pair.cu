// K4, cta_group::2 variant: a PAIR of CTAs (two SMs of one TPC, launched as a 2-CTA cluster) processes 256-row
// SUPER-tiles (tiles 2j, 2j+1 of the duration-sorted order) with ONE M=256 tcgen05.mma stream. Same phases, epilogue
// arithmetic and scratch layout as decoder_bf16.cu (see there for the algorithm and the reference lines:
// nets/modules/decoder_sa.py:577-617 loop, :146-158 prenet, :63-96 zoneout cell, :398 feat_out, :619-630 gather).
//
// Operand path:
//   * every weight stage is split along N: each CTA loads only ITS half (128 of 256 columns) and the tensor cores of
//     both SMs consume both halves, so per SM a stage is 16 KB of A + 16 KB of B instead of 16 + 32 KB -- the
//     single-CTA kernel's ring is limited by the bulk-copy bytes an SM can keep in flight (tools/umma_rate.cu);
//   * the leader CTA (cluster rank 0) issues tcgen05.mma.cta_group::2 for both; tcgen05.commit multicasts the
//     "stage free" / "accumulator ready" arrivals to both CTAs' mbarriers;
//   * the peer tells the leader "my half of stage s has landed" and "my epilogue drained accumulator b" with remote
//     mbarrier arrivals (mapa + mbarrier.arrive.relaxed.cluster) from two relay threads. .relaxed on purpose: with
//     .release.cluster every arrival cost the relay thread ~1400 cycles and serialised the ring (tools/umma_rate.cu:
//     1250-1500 vs 537 cycles per K=64 stage). The data the leader's MMA reads was written by a bulk copy (async
//     proxy) whose completion the relay thread observed on its own mbarrier; the TMEM hand-over is ordered by the
//     tcgen05 fences, not by memory ordering.
//
// TWO SUPER-TILES IN FLIGHT (p.inflight = 2). A decoder step is four dependent GEMM phases; with one tile per pair
// the tensor pipe idles at every phase boundary while the last epilogue of the previous phase runs (53 % tensor-active
// in round 1). The pair therefore runs two independent super-tiles ("slots") and alternates between them at CHUNK
// granularity: the work items of a round are  for phase: for chunk: slot 0, slot 1.  The epilogue of (slot 0, chunk c)
// runs under the MMAs of (slot 1, chunk c), and the late K-slices of a phase (the operand the previous phase's epilogue
// produces) are needed one whole chunk later than before. Accumulator buffers alternate with the global item counter
// exactly as before, so with a single active slot (tail of the list, tiny batches, inflight = 1) the stream is the
// old one. The two slots are independent tiles at independent step numbers: when a slot's tile ends it takes the next
// super-tile of the pair's list (longest-processing-time schedule, fcl_decoder_schedule). Rows keep their arithmetic:
// results are bit-identical to decoder_bf16.cu (tests/test_gpu_bf16.py, tests/test_gpu_scale.py).
#include "common.cuh"
#include "umma.cuh"

namespace fcl {
namespace pair {
using namespace umma;

constexpr int kDbThreads = 640;
constexpr int kDbStages = 6;
constexpr uint32_t kABytes = 128u * 64u * 2u;           // one A stage: 128 rows x 64 k (bf16)
constexpr uint32_t kBBytesMax = 128u * 64u * 2u;        // one B stage of THIS CTA: half of the 256 columns x 64 k
constexpr uint32_t kStageBytes = kABytes + kBBytesMax;  // 32 KB
constexpr int kEpiThreads = 512;
constexpr int kDbMaxTilesPerCta = 512;
constexpr int kSlots = 2;

struct DbShared {
  uint64_t full[kDbStages], empty[kDbStages];
  uint64_t tmem_full[2], tmem_empty[2];
  uint64_t a_ready[kSlots][4];     // per slot: x1, x2, z0', z1' operand images complete (epilogue -> producer)
  uint64_t peer_full[kDbStages];   // leader only: the peer CTA's half of stage s has landed (remote arrive)
  uint64_t peer_tmem_empty[2];     // leader only: the peer's epilogue drained accumulator buffer b (remote arrive)
  uint32_t tmem_base;
  int n_my_tiles;
  int my_tiles[kDbMaxTilesPerCta];
  int4 rowinfo[kSlots][128];       // per slot, per tile row: {row (-1 = none), duration, frame offset, utterance}
  int rowphone[kSlots][128];
};

// activation scratch (bytes). Per CTA and slot: x1 | x2 (private); z images: set `slot` of z0a z0b z1a z1b.
__host__ __device__ inline size_t db_priv_bytes(int U) { return 2 * (size_t)U * 128 * 2; }
__host__ __device__ inline size_t db_x1_off() { return 0; }
__host__ __device__ inline size_t db_x2_off(int U) { return (size_t)U * 128 * 2; }
__host__ __device__ inline size_t db_shared_bytes(int H) { return 8 * (size_t)H * 128 * 2; }
__host__ __device__ inline size_t db_z_off(int H, int set, int which /*0..3: z0a z0b z1a z1b*/) {
  return (size_t)(set * 4 + which) * H * 128 * 2;
}

// optional timeline trace of CTA 0 (debug/profiling aid; p.trace == nullptr in production).
// record = {event id, clock64}; ids: 1000*slot + 100+phase*10+chunk (MMA: accumulator free), 300+.. (MMA: chunk issued),
// 400+.. (epilogue: accumulator ready), 500+.. (epilogue: chunk done), 600+phase (producer: item start).
__device__ __forceinline__ void db_trace(const FclDecoderBf16Params& p, int id) {
  if (p.trace && blockIdx.x == 0) {
    const unsigned long long n = atomicAdd(reinterpret_cast<unsigned long long*>(p.trace), 1ull);
    if (n < (unsigned long long)p.trace_cap) {
      p.trace[2 + 2 * n] = id;
      p.trace[3 + 2 * n] = clock64();
    }
  }
}

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive (count 1) on the mbarrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[128 rows per CTA] * B[N/2 columns per CTA]^T ; issued by ONE thread of the leader CTA
__device__ __forceinline__ void mma2_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// both CTAs' mbarriers (same offset) get one arrival when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void mma2_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

// One mbarrier arrival per epilogue WARP (the barriers count 16, not 512): the lanes' writes / tcgen05.ld are ordered
// before lane 0's arrive by the warp barrier.
__device__ __forceinline__ void warp_arrive(uint64_t* bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

struct DbDims {
  int kU, kH, kE, gate_chunks;
  // byte offset of chunk c of a phase in the weight stream (blocks are [chunk][k stage])
  __device__ __forceinline__ size_t w_off(int phase, int c, uint32_t bw, uint32_t bf) const {
    const size_t l0 = (size_t)kU * bw, l1 = l0 + (size_t)gate_chunks * (kE + kH + kU) * bw;
    const size_t f = l1 + (size_t)gate_chunks * 2 * kH * bw, pc = f + (size_t)(kE + kH) * bf;
    if (phase == 0) return 0;
    if (phase == 1) return l0 + (size_t)c * (kE + kH + kU) * bw;
    if (phase == 2) return l1 + (size_t)c * 2 * kH * bw;
    return c == 0 ? f : pc;
  }
  __device__ __forceinline__ int nchunks(int phase) const { return phase == 0 ? 1 : phase == 3 ? 2 : gate_chunks; }
  __device__ __forceinline__ int kstages(int phase) const {
    return phase == 0 ? kU : phase == 1 ? kE + kH + kU : phase == 2 ? 2 * kH : kE + kH;
  }
  // first K stage (of chunk 0) that needs the operand produced by the previous phase
  __device__ __forceinline__ int late_stage(int phase) const {
    return phase == 0 ? 0 : phase == 1 ? kE + kH : phase == 2 ? kH : kE;
  }
};

// The two slots of a pair and the super-tile each one is working on. Every role (producer, MMA issuer, relay,
// epilogue) runs this same little state machine, so they all enumerate the same sequence of work items.
struct Sched {
  int st[kSlots], m[kSlots], steps[kSlots];
  int tk;
  uint32_t fresh;                                       // bit s: the slot's tile starts in this round (epilogue: tile init)
  __device__ __forceinline__ void init() {
    tk = 0; fresh = 0;
#pragma unroll
    for (int s = 0; s < kSlots; ++s) { st[s] = -1; m[s] = 0; steps[s] = 0; }
  }
  // advance to the next round: finished slots take the next super-tile of the list. false = nothing left.
  __device__ __forceinline__ bool next_round(const FclDecoderBf16Params& p, const int* list, int n_list, int n_slots) {
    bool any = false;
    fresh = 0;
#pragma unroll
    for (int s = 0; s < kSlots; ++s) {
      if (s >= n_slots) continue;
      if (st[s] >= 0 && ++m[s] >= steps[s]) st[s] = -1;
      while (st[s] < 0 && tk < n_list) {                // skip super-tiles without a single step (all durations zero)
        const int t = list[tk++];
        const int n = min(max(p.dur[p.order[(size_t)t * 256]], 0), FCL_MAX_DURATION);   // the pair runs the longer tile's steps
        if (n > 0) { st[s] = t; m[s] = 0; steps[s] = n; fresh |= 1u << s; }
      }
      any = any || st[s] >= 0;
    }
    return any;
  }
  // does (slot, phase, chunk) exist in this round? (the composed prenet chunk is skipped on a tile's last step)
  __device__ __forceinline__ bool has(int s, int phase, int c) const {
    return st[s] >= 0 && !(phase == 3 && c == 1 && m[s] + 1 == steps[s]);
  }
};

// prenet epilogue for 64 columns of one row: bias, ReLU, counter-based dropout, bf16 operand image
// (decoder_sa.py:146-158). `acc` == nullptr means a zero pre-activation (the step-0 input frame is zero).
__device__ __forceinline__ void prenet_store16(const float* v, const float* __restrict__ bias, int col0, int r, uint8_t* dst,
                                               bool use_drop, uint32_t drop_thr, float drop_scale, uint64_t seed,
                                               uint32_t utt, uint32_t ph, uint32_t step, uint32_t layer) {
#pragma unroll
  for (int h8 = 0; h8 < 2; ++h8) {                         // 8 columns share one Philox call
    Philox4 rnd = Philox4{0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
    if (use_drop) rnd = dropout_words(seed, utt, ph, step, layer, (uint32_t)((col0 >> 3) + h8));
    const uint32_t wv[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
    const float4 ba = __ldg(reinterpret_cast<const float4*>(bias + col0 + 8 * h8));
    const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + col0 + 8 * h8) + 1);
    const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t u16 = (j & 1) ? (wv[j >> 1] >> 16) : (wv[j >> 1] & 0xFFFFu);
      const float y = fmaxf((v ? v[8 * h8 + j] : 0.f) + bv[j], 0.f) * drop_scale;
      x[j] = u16 >= drop_thr ? y : 0.f;
    }
    uint4 w;
    w.x = pack_op(x[0], x[1]); w.y = pack_op(x[2], x[3]);
    w.z = pack_op(x[4], x[5]); w.w = pack_op(x[6], x[7]);
    *reinterpret_cast<uint4*>(dst + ((size_t)((col0 >> 3) + h8) * 128 + r) * 16) = w;
  }
}

__global__ void __launch_bounds__(kDbThreads, 1)
decoder_bf16_pair_kernel(FclDecoderBf16Params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ DbShared sh;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = p.dunits, U = p.prenet_units, O = p.odim, E = p.eunits;
  DbDims dm;
  dm.kU = U / 64; dm.kH = H / 64; dm.kE = E / 64; dm.gate_chunks = 4 * H / 256;
  const uint32_t rank = cluster_ctarank();              // 0 = leader (issues the MMAs), 1 = peer
  const int grp = (int)blockIdx.x >> 1;                 // the pair = one schedule slot; it walks SUPER-tiles (2 tiles)
  const int inflight = p.inflight >= 2 ? 2 : 1;
  // scratch of this CTA: slot s uses x1|x2 block (2*cta + s), z image set s, cell-state block (2*cta + s)
  uint8_t* act_base = reinterpret_cast<uint8_t*>(p.act_priv) + (size_t)blockIdx.x * 2 * db_priv_bytes(U);
  uint8_t* zsh = reinterpret_cast<uint8_t*>(p.act_shared) + (size_t)blockIdx.x * db_shared_bytes(H);
  float* cws_base = p.c_ws + (size_t)blockIdx.x * 2 * 2 * H * 128;

  // this CTA's tile list (longest-processing-time schedule)
  if (tid == 0) sh.n_my_tiles = 0;
  __syncthreads();
  const int n_super = (p.n_tiles + 1) >> 1;
  for (int t = tid; t < n_super; t += kDbThreads) {
    if (p.tile_slot[t] == grp) {
      const int k = p.tile_rank[t];
      if (k < kDbMaxTilesPerCta) { sh.my_tiles[k] = t; atomicMax(&sh.n_my_tiles, k + 1); }
    }
  }
  if (tid == 0) {
    for (int s = 0; s < kDbStages; ++s) { mbar_init(&sh.full[s], 1); mbar_init(&sh.empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&sh.tmem_full[b], 1); mbar_init(&sh.tmem_empty[b], kEpiThreads / 32); }
    for (int s = 0; s < kSlots; ++s)
      for (int i = 0; i < 4; ++i) mbar_init(&sh.a_ready[s][i], kEpiThreads / 32);
    for (int s = 0; s < kDbStages; ++s) mbar_init(&sh.peer_full[s], 1);
    for (int b = 0; b < 2; ++b) mbar_init(&sh.peer_tmem_empty[b], 1);
    fence_barrier_init();
  }
  cluster_sync_all();                                   // barriers of both CTAs initialised before any remote arrive
  if (warp == 2) tmem_alloc2(&sh.tmem_base, 512);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = sh.tmem_base;
  // per-CTA halves: 128 of the 256 gate/prenet columns, 64 of the feat_out columns (odim zero-padded to 128)
  const uint32_t b_bytes_wide = 128u * 64u * 2u, b_bytes_feat = 64u * 64u * 2u;
  Sched sc;
  sc.init();
  const int n_list = sh.n_my_tiles;

  // Register budget (see decoder_bf16.cu): warps 0-3 give up 32 registers per thread, the epilogue threads get 16 more.
  // Each setmaxnreg has to dominate the code it is meant for, hence the two-level role dispatch.
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
  if (warp == 0) {
    // ================================================================ producer
    if (elect_one()) {
      uint32_t stage = 0, sphase = 0;                 // ring position / parity
      uint32_t rdy[kSlots][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};   // parity of each a_ready barrier
      while (sc.next_round(p, sh.my_tiles, n_list, inflight)) {
        for (int phase = 0; phase < 4; ++phase) {
          const int kst = dm.kstages(phase), late = dm.late_stage(phase);
          for (int c = 0; c < dm.nchunks(phase); ++c) {
#pragma unroll
            for (int s = 0; s < kSlots; ++s) {
              if (!sc.has(s, phase, c)) continue;
              if (tid == 0) db_trace(p, 1000 * s + 600 + phase);
              const int tile = min(2 * sc.st[s] + (int)rank, p.n_tiles - 1);    // an odd tile count leaves the last peer a dummy (masked) tile
              const uint8_t* himg = reinterpret_cast<const uint8_t*>(p.hn_img) + (size_t)tile * E * 128 * 2;
              const uint8_t* act = act_base + (size_t)s * db_priv_bytes(U);
              const int zp = sc.m[s] & 1;
              const uint8_t* z0cur = zsh + db_z_off(H, s, zp), *z0new = zsh + db_z_off(H, s, zp ^ 1);
              const uint8_t* z1cur = zsh + db_z_off(H, s, 2 + zp), *z1new = zsh + db_z_off(H, s, 2 + (zp ^ 1));
              const uint32_t bb = (phase == 3 && c == 0) ? b_bytes_feat : b_bytes_wide;
              // stage blocks hold both halves back to back: [rank 0 half][rank 1 half]
              const uint8_t* wptr = reinterpret_cast<const uint8_t*>(p.w_stream) +
                                    dm.w_off(phase, c, 2 * b_bytes_wide, 2 * b_bytes_feat) + (size_t)rank * bb;
              for (int ks = 0; ks < kst; ++ks) {
                if (c == 0 && ks == late) {            // operand written by this slot's previous-phase epilogue
                  mbar_wait(&sh.a_ready[s][phase], rdy[s][phase]);
                  rdy[s][phase] ^= 1u;
                }
                const uint8_t* asrc;
                if (phase == 0) asrc = act + db_x1_off() + (size_t)ks * kABytes;
                else if (phase == 1) asrc = ks < dm.kE ? himg + (size_t)ks * kABytes
                                          : ks < dm.kE + dm.kH ? z0cur + (size_t)(ks - dm.kE) * kABytes
                                                               : act + db_x2_off(U) + (size_t)(ks - dm.kE - dm.kH) * kABytes;
                else if (phase == 2) asrc = ks < dm.kH ? z1cur + (size_t)ks * kABytes : z0new + (size_t)(ks - dm.kH) * kABytes;
                else asrc = ks < dm.kE ? himg + (size_t)ks * kABytes : z1new + (size_t)(ks - dm.kE) * kABytes;
                mbar_wait(&sh.empty[stage], sphase ^ 1u);
                mbar_arrive_expect_tx(&sh.full[stage], kABytes + bb);
                uint8_t* stg = smem + (size_t)stage * kStageBytes;
                bulk_g2s(stg, asrc, kABytes, &sh.full[stage]);
                bulk_g2s(stg + kABytes, wptr, bb, &sh.full[stage]);
                wptr += 2 * bb;
                if (++stage == kDbStages) { stage = 0; sphase ^= 1u; }
              }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================================================ MMA issuer (leader) / stage-full relay (peer)
    if (elect_one()) {
      uint32_t stage = 0, sphase = 0;
      uint32_t chunk_ctr = 0;                          // accumulator buffer = chunk_ctr & 1
      const uint32_t idesc_wide = idesc_op_f32(256u, 256u), idesc_feat = idesc_op_f32(256u, 128u);
      // descriptors built incrementally (see decoder_bf16.cu): low word = (address >> 4) | (LBO >> 4) << 16
      const uint32_t ring_lo = smem_u32(smem) >> 4;
      constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);
      constexpr uint32_t kALo = (2048u >> 4) << 16;
      uint32_t s_lo = ring_lo;
      while (sc.next_round(p, sh.my_tiles, n_list, inflight)) {
        for (int phase = 0; phase < 4; ++phase) {
          const int kst = dm.kstages(phase);
          for (int c = 0; c < dm.nchunks(phase); ++c) {
#pragma unroll
            for (int s = 0; s < kSlots; ++s) {
              if (!sc.has(s, phase, c)) continue;
              const bool feat = phase == 3 && c == 0;
              const uint32_t idesc = feat ? idesc_feat : idesc_wide;
              const uint32_t b_lbo = feat ? 64u * 16u : 128u * 16u;          // rows of THIS CTA's half x 16 B
              const uint32_t b_lo = (kABytes >> 4) + ((b_lbo >> 4) << 16), b_kstep = (2u * b_lbo) >> 4;
              const uint32_t buf = chunk_ctr & 1u, use = chunk_ctr >> 1;
              if (rank == 0) {
                mbar_wait(&sh.tmem_empty[buf], (use & 1u) ^ 1u);
                if (use > 0) mbar_wait(&sh.peer_tmem_empty[buf], (use & 1u) ^ 1u);   // completion #(use-1): the peer drained it too
                tc_fence_after();
                db_trace(p, 1000 * s + 100 + phase * 10 + c);
              }
              const uint32_t d_tmem = tmem + buf * 256u;
              for (int ks = 0; ks < kst; ++ks) {
                mbar_wait(&sh.full[stage], sphase);
                if (rank == 1) {
                  mbar_arrive_remote(&sh.peer_full[stage], 0);               // tell the leader our half has landed
                } else {
                  mbar_wait(&sh.peer_full[stage], sphase);
                  tc_fence_after();
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    const uint64_t ad = ((uint64_t)kDescHi << 32) | (s_lo + (uint32_t)k * (4096u >> 4) + kALo);
                    const uint64_t bd = ((uint64_t)kDescHi << 32) | (s_lo + b_lo + (uint32_t)k * b_kstep);
                    mma2_bf16_ss(d_tmem, ad, bd, idesc, (ks > 0 || k > 0) ? 1u : 0u);
                  }
                  mma2_commit(&sh.empty[stage]);                             // frees the stage in BOTH CTAs
                }
                s_lo += kStageBytes >> 4;
                if (++stage == kDbStages) { stage = 0; sphase ^= 1u; s_lo = ring_lo; }
              }
              if (rank == 0) {
                mma2_commit(&sh.tmem_full[buf]);                             // accumulator ready in BOTH CTAs
                db_trace(p, 1000 * s + 300 + phase * 10 + c);
              }
              ++chunk_ctr;
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    // ================================================================ peer only: "accumulator drained" relay
    if (rank == 1 && elect_one()) {
      uint32_t chunk_ctr = 0;
      while (sc.next_round(p, sh.my_tiles, n_list, inflight)) {
        for (int phase = 0; phase < 4; ++phase) {
          for (int c = 0; c < dm.nchunks(phase); ++c) {
#pragma unroll
            for (int s = 0; s < kSlots; ++s) {
              if (!sc.has(s, phase, c)) continue;
              const uint32_t buf = chunk_ctr & 1u, use = chunk_ctr >> 1;
              if (use > 0) {
                mbar_wait(&sh.tmem_empty[buf], (use & 1u) ^ 1u);             // our epilogue finished use-1 of this buffer
                mbar_arrive_remote(&sh.peer_tmem_empty[buf], 0);
              }
              ++chunk_ctr;
            }
          }
        }
      }
    }
    __syncwarp();
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    // ================================================================ epilogue (512 threads)
    const int q = warp & 3, cs = (warp - 4) >> 2;      // TMEM lane quarter, column quarter
    const int r = q * 32 + lane;                       // row within the tile == TMEM lane
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
    uint32_t chunk_ctr = 0;
    const float zo = p.zoneout, zk = 1.0f - p.zoneout;
    const bool use_drop = p.dropout_p > 0.f;
    const uint32_t drop_thr = dropout_threshold16(p.dropout_p);
    const float drop_scale = use_drop ? 1.0f / (1.0f - p.dropout_p) : 1.0f;

    while (sc.next_round(p, sh.my_tiles, n_list, inflight)) {
      // ---- tile init of the slots that start a tile in this round: row table, x1 of step 0 (the first input frame
      // is zero: prenet.0 sees only its bias) and zero z images
      if (sc.fresh) epi_bar_sync();   // nobody still reads the row table of the tile that just ended
#pragma unroll 1
      for (int s = 0; s < kSlots; ++s) {
        if (!((sc.fresh >> s) & 1u)) continue;
        const int tile = 2 * sc.st[s] + (int)rank;
        const int sidx = tile * 128 + r;
        int row = -1, d = 0, foff = 0, utt = 0, ph = 0;
        if (tile < p.n_tiles && sidx < p.n_rows) {
          row = p.order[sidx];
          d = min(max(p.dur[row], 0), FCL_MAX_DURATION);
          foff = p.frame_off[row];
          utt = p.row_utt[row];
          ph = p.row_phone[row];
        }
        if (cs == 0) { sh.rowinfo[s][r] = make_int4(row, d, foff, utt); sh.rowphone[s][r] = ph; }
        uint8_t* act = act_base + (size_t)s * db_priv_bytes(U);
#pragma unroll 1
        for (int g = 0; g < 4; ++g)
          prenet_store16(nullptr, p.bp0, cs * 64 + g * 16, r, act + db_x1_off(), use_drop, drop_thr, drop_scale,
                         p.dropout_seed, (uint32_t)utt, (uint32_t)ph, 0u, 0u);
        const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
        for (int kc = cs; kc < H / 8; kc += 4) {
          *reinterpret_cast<uint4*>(zsh + db_z_off(H, s, 0) + ((size_t)kc * 128 + r) * 16) = z4;
          *reinterpret_cast<uint4*>(zsh + db_z_off(H, s, 2) + ((size_t)kc * 128 + r) * 16) = z4;
        }
        fence_proxy_async_global();
        warp_arrive(&sh.a_ready[s][0], lane);
      }
      if (sc.fresh) epi_bar_sync();   // the row tables are read by all four column quarters

      for (int phase = 0; phase < 4; ++phase) {
#pragma unroll 1
        for (int c = 0; c < dm.nchunks(phase); ++c) {
#pragma unroll 1
          for (int s = 0; s < kSlots; ++s) {
            if (!sc.has(s, phase, c)) continue;
            const int m = sc.m[s];
            uint8_t* act = act_base + (size_t)s * db_priv_bytes(U);
            const uint32_t buf = chunk_ctr & 1u, use = chunk_ctr >> 1;
            if (phase == 0) {
              // ---------------- P1: prenet layer 1 (bias, ReLU, dropout) -> x2 image; 64 columns per thread
              const int4 ri = sh.rowinfo[s][r];
              const int ph = sh.rowphone[s][r];
              mbar_wait(&sh.tmem_full[buf], use & 1u);
              tc_fence_after();
              if (tid == 128) db_trace(p, 1000 * s + 400);
#pragma unroll 1
              for (int g = 0; g < 4; ++g) {
                float v[16];
                const int col0 = cs * 64 + g * 16;
                tmem_ld16(lane_addr + buf * 256u + (uint32_t)col0, v);
                prenet_store16(v, p.bp1, col0, r, act + db_x2_off(U), use_drop, drop_thr, drop_scale, p.dropout_seed,
                               (uint32_t)ri.w, (uint32_t)ph, (uint32_t)m, 1u);
              }
              tc_fence_before();
              warp_arrive(&sh.tmem_empty[buf], lane);
              ++chunk_ctr;
              fence_proxy_async_global();
              warp_arrive(&sh.a_ready[s][1], lane);
              if (tid == 128) db_trace(p, 1000 * s + 500);
            } else if (phase == 1 || phase == 2) {
              // ---------------- L0, L1: zoneout LSTM cells; per chunk this thread owns 16 hidden units of its row.
              const int layer = phase - 1;
              const int zp = m & 1;
              const uint8_t* zcur = zsh + db_z_off(H, s, 2 * layer + zp);
              uint8_t* znew = zsh + db_z_off(H, s, 2 * layer + (zp ^ 1));
              float* cl = cws_base + ((size_t)s * 2 + layer) * H * 128;
              const float* bias = layer == 0 ? p.b0 : p.b1;
              float pos = 0.f;
              if (layer == 0) {
                const int4 ri = sh.rowinfo[s][r];
                pos = (ri.x >= 0 && m < ri.y) ? __fdiv_rn((float)m, (float)ri.y) : 0.f;
              }
              float c_cur[16];
              uint4 z_cur[2];
              const int u0 = c * 64 + cs * 16;                          // first of this thread's 16 hidden units
              // old cell state / old z of this chunk: requested BEFORE waiting for the accumulator (their L2 latency hides
              // behind the MMAs)
#pragma unroll
              for (int j = 0; j < 16; ++j) c_cur[j] = m == 0 ? 0.f : __ldcg(cl + (size_t)(u0 + j) * 128 + r);
              z_cur[0] = __ldcg(reinterpret_cast<const uint4*>(zcur + ((size_t)(u0 >> 3) * 128 + r) * 16));
              z_cur[1] = __ldcg(reinterpret_cast<const uint4*>(zcur + ((size_t)((u0 >> 3) + 1) * 128 + r) * 16));
              mbar_wait(&sh.tmem_full[buf], use & 1u);
              tc_fence_after();
              if (tid == 128) db_trace(p, 1000 * s + 400 + phase * 10 + c);
              uint32_t zout[8];
#pragma unroll
              for (int g = 0; g < 4; ++g) {                             // 4 units (16 accumulator columns) at a time
                float v[16];
                tmem_ld16(lane_addr + buf * 256u + (uint32_t)(cs * 64 + g * 16), v);
                float zn[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const int ul = g * 4 + j, u = u0 + ul;
                  float4 add = __ldg(reinterpret_cast<const float4*>(bias + 4 * u));
                  if (layer == 0) {
                    const float4 wp = __ldg(reinterpret_cast<const float4*>(p.wpos + 4 * u));
                    add.x = fmaf(pos, wp.x, add.x); add.y = fmaf(pos, wp.y, add.y);
                    add.z = fmaf(pos, wp.z, add.z); add.w = fmaf(pos, wp.w, add.w);
                  }
                  const float ig = sigmoid_fast(v[4 * j] + add.x), fg = sigmoid_fast(v[4 * j + 1] + add.y);
                  const float gg = tanh_fast(v[4 * j + 2] + add.z), og = sigmoid_fast(v[4 * j + 3] + add.w);
                  const float cold = c_cur[ul];
                  const float cn = fmaf(fg, cold, ig * gg);
                  const float hn = og * tanh_fast(cn);
                  const uint4 zq = z_cur[ul >> 3];
                  const uint32_t zw = ((ul >> 1) & 3) == 0 ? zq.x : ((ul >> 1) & 3) == 1 ? zq.y : ((ul >> 1) & 3) == 2 ? zq.z : zq.w;
                  const float zold = (ul & 1) ? op_hi(zw) : op_lo(zw);
                  zn[j] = fmaf(zo, zold, zk * hn);                      // decoder_sa.py:95-96 (eval blend)
                  cl[(size_t)u * 128 + r] = fmaf(zo, cold, zk * cn);
                }
                zout[2 * g] = pack_op(zn[0], zn[1]);
                zout[2 * g + 1] = pack_op(zn[2], zn[3]);
              }
              tc_fence_before();
              warp_arrive(&sh.tmem_empty[buf], lane);
              ++chunk_ctr;
              *reinterpret_cast<uint4*>(znew + ((size_t)(u0 >> 3) * 128 + r) * 16) = make_uint4(zout[0], zout[1], zout[2], zout[3]);
              *reinterpret_cast<uint4*>(znew + ((size_t)((u0 >> 3) + 1) * 128 + r) * 16) = make_uint4(zout[4], zout[5], zout[6], zout[7]);
              if (tid == 128) db_trace(p, 1000 * s + 500 + phase * 10 + c);
              if (c == dm.gate_chunks - 1) {                            // the whole z' image of this slot is written
                fence_proxy_async_global();
                warp_arrive(&sh.a_ready[s][2 + layer], lane);
              }
            } else if (c == 0) {
              // ---------------- FP chunk 0: feat_out -> output frame, stored straight to its final (ragged) position
              const int4 ri = sh.rowinfo[s][r];
              mbar_wait(&sh.tmem_full[buf], use & 1u);
              tc_fence_after();
              if (tid == 128) db_trace(p, 1000 * s + 430);
              for (int g = cs; g < O / 16; g += 4) {                     // 16-column groups dealt over the 4 column sets
                float v[16];
                tmem_ld16(lane_addr + buf * 256u + (uint32_t)(g * 16), v);
                if (ri.x >= 0 && m < ri.y) {                             // exhausted rows are masked (decoder_sa.py:625-629)
                  float4* o = reinterpret_cast<float4*>(p.before + ((size_t)ri.z + m) * O + g * 16);
#pragma unroll
                  for (int qd = 0; qd < 4; ++qd)                          // streaming store: written once, read by the next kernel
                    __stcs(o + qd, make_float4(v[4 * qd], v[4 * qd + 1], v[4 * qd + 2], v[4 * qd + 3]));
                }
              }
              tc_fence_before();
              warp_arrive(&sh.tmem_empty[buf], lane);
              ++chunk_ctr;
              if (tid == 128) db_trace(p, 1000 * s + 530);
            } else {
              // ---------------- FP chunk 1: prenet layer 0 of the NEXT step (composed with feat_out) -> x1 image
              const int4 ri = sh.rowinfo[s][r];
              const int ph = sh.rowphone[s][r];
              mbar_wait(&sh.tmem_full[buf], use & 1u);
              tc_fence_after();
              if (tid == 128) db_trace(p, 1000 * s + 431);
#pragma unroll 1
              for (int g = 0; g < 4; ++g) {
                float v[16];
                const int col0 = cs * 64 + g * 16;
                tmem_ld16(lane_addr + buf * 256u + (uint32_t)col0, v);
                prenet_store16(v, p.bp0, col0, r, act + db_x1_off(), use_drop, drop_thr, drop_scale, p.dropout_seed,
                               (uint32_t)ri.w, (uint32_t)ph, (uint32_t)(m + 1), 0u);
              }
              tc_fence_before();
              warp_arrive(&sh.tmem_empty[buf], lane);
              ++chunk_ctr;
              fence_proxy_async_global();
              warp_arrive(&sh.a_ready[s][0], lane);
              if (tid == 128) db_trace(p, 1000 * s + 531);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();                                   // the leader's MMAs read the peer's shared memory: leave together
  if (warp == 2) tmem_dealloc2(tmem, 512);
}

}  // namespace pair


}  // namespace fcl

extern "C" int fcl_decoder_bf16_pair(const FclDecoderBf16Params* p, void* stream) {
  using namespace fcl;
  using namespace fcl::pair;
  FCL_REQUIRE(p && p->order && p->dur && p->frame_off && p->row_utt && p->row_phone && p->hn_img &&
                  p->w_stream && p->bp0 && p->bp1 && p->wpos && p->b0 && p->b1 && p->act_priv && p->act_shared &&
                  p->c_ws && p->before && p->tile_slot && p->tile_rank,
              "null pointer");
  FCL_REQUIRE(p->eunits % 64 == 0 && p->eunits >= 64, "eunits must be a multiple of 64");
  FCL_REQUIRE(p->n_rows > 0 && p->n_tiles == (p->n_rows + 127) / 128, "n_tiles must be ceil(n_rows / 128)");
  FCL_REQUIRE(p->prenet_units == 256, "prenet_units must be 256 (one 256-column chunk)");
  FCL_REQUIRE(p->dunits % 64 == 0 && p->dunits >= 64, "dunits must be a multiple of 64");
  FCL_REQUIRE(p->odim % 16 == 0 && p->odim <= 128, "odim must be a multiple of 16, <= 128");
  FCL_REQUIRE(p->n_slots >= 2 && p->n_slots % 2 == 0, "n_slots must be even (CTA pairs)");
  FCL_REQUIRE(p->zoneout >= 0.f && p->zoneout < 1.f && p->dropout_p >= 0.f && p->dropout_p < 1.f, "bad rates");
  FCL_REQUIRE((long long)((p->n_tiles + 1) / 2) <= (long long)kDbMaxTilesPerCta * (p->n_slots / 2), "too many tiles");
  FCL_REQUIRE(!p->tf_x1, "teacher forcing is implemented in fcl_decoder_bf16 and fcl_decoder_bf16_pair_v1");
  const size_t smem = (size_t)kDbStages * kStageBytes;
  if (int rc = ensure_dyn_smem(decoder_bf16_pair_kernel, smem, "fcl_decoder_bf16_pair")) return rc;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)p->n_slots);
  cfg.blockDim = dim3(kDbThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = as_stream(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, decoder_bf16_pair_kernel, *p);
  if (e != cudaSuccess) { set_error("fcl_decoder_bf16_pair: %s", cudaGetErrorString(e)); return FCL_ECUDA; }
  return check_launch("fcl_decoder_bf16_pair");
}
