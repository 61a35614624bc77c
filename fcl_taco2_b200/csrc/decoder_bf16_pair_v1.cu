// The cta_group::2 decoder the engine runs for large batches (one super-tile per CTA pair at a time, per-role tile
// loops). decoder_bf16_pair.cu is the slot-scheduler variant with two super-tiles in flight (measured slower).
// K4, cta_group::2 variant: a PAIR of CTAs (two SMs of one TPC, launched as a 2-CTA cluster) processes two
// 128-row tiles with ONE M=256 tcgen05.mma stream. Same phases, epilogue and scratch layout as decoder_bf16.cu
// (see there for the algorithm and the reference lines); what changes is the operand path:
//   * every weight stage is split along N: each CTA loads only ITS half (128 of 256 columns) and the tensor cores
//     of both SMs consume both halves, so per SM a stage is 16 KB of A + 16 KB of B instead of 16 + 32 KB -- the
//     single-CTA kernel's ring is limited by the bulk-copy bytes an SM can keep in flight (a copy takes ~1080
//     cycles whatever its size; 48 KB per 512-cycle stage does not fit, 32 KB does: tools/umma_rate.cu);
//   * the leader CTA (cluster rank 0) issues tcgen05.mma.cta_group::2 for both; tcgen05.commit multicasts the
//     "stage free" / "accumulator ready" arrivals to both CTAs' mbarriers;
//   * the peer tells the leader "my half of stage s has landed" and "my epilogue drained accumulator b" with remote
//     mbarrier arrivals (mapa + mbarrier.arrive.relaxed.cluster) from two relay threads. They must be .relaxed: with
//     .release.cluster every arrival cost the relay thread ~1400 cycles and serialised the ring (tools/umma_rate.cu
//     reproduces it in isolation: 1250-1500 vs 537 cycles per K=64 stage). Ordering does not need the release: the
//     bytes were written by the bulk copy whose completion the relay thread observed (acquire) before it signals.
// The pair walks SUPER-tiles (tiles 2j, 2j+1 of the duration-sorted order) for the step count of the longer one.
//
// STATUS (round 1): bit-identical to decoder_bf16.cu (tests/test_gpu_bf16.py::test_decoder_pair_mode_bit_identical) and
// faster once every SM has a tile anyway: S batch 1024 2.13 ms vs 2.40 ms, T batch 1024 15.4 vs 18.8 ms (the MMA
// stream runs at 500-570 cycles per K stage against ~800 for the single-CTA kernel, whose ring is limited by the
// bulk-copy bytes in flight per SM). Engine.use_pair = None picks it when n_tiles >= SM count.
// WHAT BOUNDS IT (round 2, profiles/r02_decoder_whatif.md; tools/decoder_prof.py, tools/decoder_whatif.py with the
// -DFCL_DEC_PROF build): the EPILOGUE is the longer pole, not the operand stream. Per tile-step the epilogue thread is
// busy ~74 k cycles (P1 7.5 k, 4 x 7.6 k cell 0, 4 x 6.5 k cell 1, feat_out 2.7 k, composed prenet.0 7.3 k) against ~51 k
// cycles of MMAs; with a skeleton epilogue (wait, hand over) the same kernel takes 1.16 ms instead of 1.81 ms. Removing
// all A-operand copies buys 6 %, the cell-state loads/stores 12 %, the image stores 6 %, Philox 4 %, the cell math 7 %.
// Tried this round and measured no better: operand images resident in shared memory (-20 % L2 -> SM bytes, ring of 3:
// 1.85 ms), K = 128 ring items (1.86 ms), L2 evict_last / prefetch for the cell state (no change; evict_last lines slowed
// the postnet that follows by up to 1 ms), re-loading each group's registers for the next chunk (spills: 2.19 ms).
// Kept: the epilogue constants in shared memory (-1..3 %), the cell-state stores after the phase hand-over.
// Tried and rejected (measured slower, both kernels): evaluating the dropout Philox stream ahead of the prenet
// epilogues, either in the epilogue warps' waiting windows or in two extra warps through shared memory
// (S batch 1024: 2.40 -> 2.64 ms single, 2.13 -> 2.48 ms pair).
#include "common.cuh"
#include "umma.cuh"

namespace fcl {
namespace pair_v1 {
using namespace umma;

constexpr int kDbThreads = 640;
constexpr int kDbStages = 6;
constexpr uint32_t kABytes = 128u * 64u * 2u;           // one A stage: 128 rows x 64 k (bf16)
constexpr uint32_t kBBytesMax = 128u * 64u * 2u;        // one B stage of THIS CTA: half of the 256 columns x 64 k
constexpr uint32_t kStageBytes = kABytes + kBBytesMax;  // 32 KB
constexpr int kEpiThreads = 512;
constexpr int kDbMaxTilesPerCta = 512;

// kernel-side parameters: the C-ABI struct plus launcher-derived switches
struct FclDecoderBf16ParamsEx : FclDecoderBf16Params {
  int smem_consts;            // epilogue constants staged in shared memory behind the ring (they fit)
};

struct DbShared {
  uint64_t full[kDbStages], empty[kDbStages];
  uint64_t tmem_full[2], tmem_empty[2];
  uint64_t a_ready[4];        // x1, x2, z0', z1' operand images complete (epilogue -> producer)
  uint64_t peer_full[kDbStages];   // leader only: the peer CTA's half of stage s has landed (remote arrive)
  uint64_t peer_tmem_empty[2];     // leader only: the peer's epilogue drained accumulator buffer b (remote arrive)
  uint32_t tmem_base;
#ifdef FCL_DEC_PROF
  uint32_t prof[48];          // counter-mode profile of CTA 0 (p.trace != nullptr && p.trace_cap < 0), see DB_PROF
#endif
  int n_my_tiles;
  int my_tiles[kDbMaxTilesPerCta];
};

// activation scratch (bytes). Private to a CTA: x1 | x2. Shared by the CTAs of a group (which split the gate
// columns of one tile): two sets (even / odd tile of the group) of z0a z0b z1a z1b.
__host__ __device__ inline size_t db_priv_bytes(int U) { return 2 * (size_t)U * 128 * 2; }
__host__ __device__ inline size_t db_x1_off() { return 0; }
__host__ __device__ inline size_t db_x2_off(int U) { return (size_t)U * 128 * 2; }
__host__ __device__ inline size_t db_shared_bytes(int H) { return 8 * (size_t)H * 128 * 2; }
__host__ __device__ inline size_t db_z_off(int H, int set, int which /*0..3: z0a z0b z1a z1b*/) {
  return (size_t)(set * 4 + which) * H * 128 * 2;
}

// tile visited at round `tk` by this CTA: from the LPT schedule (fcl_decoder_schedule) staged in shared memory.
#define DB_TILE(tk) ((tk) < sh.n_my_tiles ? sh.my_tiles[tk] : -1)

// optional timeline trace of CTA 0 (debug/profiling aid; p.trace == nullptr in production).
// record = {event id, clock64}; ids: 100+phase*10+chunk (MMA: accumulator free), 200+.. (MMA: first stage landed),
// 300+.. (MMA: chunk issued), 400+.. (epilogue: accumulator ready), 500+.. (epilogue: chunk done), 600+phase
// (producer: phase start). Phases: 0 = P1, 1 = L0, 2 = L1, 3 = FP.
__device__ __forceinline__ void db_trace(const FclDecoderBf16Params& p, int id) {
  if (p.trace && p.trace_cap > 0 && blockIdx.x == 0) {
    const unsigned long long n = atomicAdd(reinterpret_cast<unsigned long long*>(p.trace), 1ull);
    if (n < (unsigned long long)p.trace_cap) {
      p.trace[2 + 2 * n] = id;
      p.trace[3 + 2 * n] = clock64();
    }
  }
}

// Counter-mode profile (p.trace != nullptr, p.trace_cap < 0): the single-thread roles and one epilogue thread of CTA 0
// accumulate %clock deltas in shared memory (no global traffic, no atomics: ~30 cycles per probe) and CTA 0 dumps the
// 48 counters to p.trace at the end. Layout (cycles, per phase ph = 0..3 = P1, L0, L1, FP):
//   [4*ph + 0] MMA issuer: waiting for a free accumulator (epilogue-bound)   [4*ph + 1] waiting for the FIRST ring item of a chunk
//   [4*ph + 2] waiting for the other ring items (operand-stream-bound)       [4*ph + 3] issuing MMAs + commits
//   [16 + 2*ph] epilogue thread 128: waiting for the accumulator             [17 + 2*ph] epilogue body
//   [24 + ph] producer: waiting for a_ready (operand of the previous phase)  [28 + ph] producer: waiting for a free ring slot
//   [32] tile-steps of this CTA   [33] total cycles of the issuer            [34 + ph] ring items issued
//   [44] epilogue: waiting for the composed prenet.0 accumulator (FP chunk 1) [45] its body   [46], [47] residue (tile init, loop overhead)
__device__ __forceinline__ uint32_t clk32() { uint32_t c; asm volatile("mov.u32 %0, %%clock;" : "=r"(c)); return c; }
#ifdef FCL_DEC_PROF
#define DB_PROF(idx) do { if (prof) { const uint32_t n_ = clk32(); sh.prof[idx] += n_ - pt; pt = n_; } } while (0)
#define DB_PROF_ON(p) ((p).trace && (p).trace_cap < 0 && blockIdx.x == 0)
// what-if switches of the profiling build (results are garbage; they tell which traffic / work bounds the kernel):
// p.inflight bit 8: no A copies in ring items, bit 9: no cell-state loads/stores, bit 10: no operand-image stores,
// bit 11: no Philox in the prenet epilogues, bits 12-14: ring depth override (v1),
// bit 15: skeleton epilogue (wait, hand over, nothing else), bit 16: no cell math (TMEM loads + stores only),
// bit 18: skeleton + ~800 dependent integer multiply-adds per thread and chunk (issue pressure without memory traffic),
// bit 19: skeleton + the same duration asleep (3200 cycles per chunk, no issue pressure)
#define DB_WHATIF(p, bit) ((((p).inflight) >> (bit)) & 1)
#else
#define DB_PROF(idx) do { (void)prof; (void)pt; } while (0)
#define DB_PROF_ON(p) false
#define DB_WHATIF(p, bit) 0
#endif

#ifdef FCL_DEC_PROF
__device__ __noinline__ uint32_t db_dummy_alu(uint32_t x) {
  uint32_t a = x, b = x ^ 0x9E3779B9u, c = x + 0x7F4A7C15u, d = x * 3u;
#pragma unroll 1
  for (int i = 0; i < 200; ++i) { a = a * 1664525u + b; b = b * 22695477u + c; c = c * 1103515245u + d; d = d * 214013u + a; }
  return a ^ b ^ c ^ d;
}
__device__ __forceinline__ void db_dummy_work(const FclDecoderBf16Params& p, uint32_t seed) {
  if (DB_WHATIF(p, 18)) {
    const uint32_t v = db_dummy_alu(seed);
    if (v == 0x12345678u && p.trace) p.trace[60] = v;             // never true in practice: keeps the loop alive
  }
  if (DB_WHATIF(p, 19)) {
    const uint32_t t0 = clk32();
    while (clk32() - t0 < 3200u) __nanosleep(200);
  }
}
#else
__device__ __forceinline__ void db_dummy_work(const FclDecoderBf16Params&, uint32_t) {}
#endif

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
// before lane 0's arrive by the warp barrier; 16 arrivals per hand-over instead of 512 (pair kernel, S batch 1024: -1.6 %).
__device__ __forceinline__ void warp_arrive(uint64_t* bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

struct DbDims {
  int kU, kH, kE, gate_chunks;
  int C, cr;                       // CTAs cooperating on a tile, rank of this CTA among them
  // which chunks of a phase this CTA computes: the gate chunks are dealt round-robin over the group, the small
  // prenet phases are computed redundantly by everyone (their outputs stay private), feat_out by rank 0 only
  __device__ __forceinline__ bool owns(int phase, int c) const {
    if (phase == 1 || phase == 2) return (c % C) == cr;
    if (phase == 3 && c == 0) return cr == 0;
    return true;
  }
  // byte offset of chunk c of a phase in the weight stream (blocks are [chunk][k stage])
  __device__ __forceinline__ size_t w_off(int phase, int c, uint32_t bw, uint32_t bf) const {
    const size_t l0 = (size_t)kU * bw, l1 = l0 + (size_t)gate_chunks * (kE + kH + kU) * bw;
    const size_t f = l1 + (size_t)gate_chunks * 2 * kH * bw, pc = f + (size_t)(kE + kH) * bf;
    if (phase == 0) return 0;
    if (phase == 1) return l0 + (size_t)c * (kE + kH + kU) * bw;
    if (phase == 2) return l1 + (size_t)c * 2 * kH * bw;
    return c == 0 ? f : pc;
  }
  __device__ __forceinline__ int nchunks(int phase, bool last_step) const {
    return phase == 0 ? 1 : phase == 3 ? (last_step ? 1 : 2) : gate_chunks;
  }
  __device__ __forceinline__ int kstages(int phase) const {
    return phase == 0 ? kU : phase == 1 ? kE + kH + kU : phase == 2 ? 2 * kH : kE + kH;
  }
  // first K stage (of chunk 0) that needs the operand produced by the previous phase
  __device__ __forceinline__ int late_stage(int phase) const {
    return phase == 0 ? 0 : phase == 1 ? kE + kH : phase == 2 ? kH : kE;
  }
};

// prenet epilogue for 64 columns of one row: bias, ReLU, counter-based dropout, bf16 operand image
// (decoder_sa.py:146-158). `acc` == nullptr means a zero pre-activation (the step-0 input frame is zero).
__device__ __forceinline__ void prenet_store16(const float* v, const float* bias, int col0, int r, uint8_t* dst,
                                               bool use_drop, uint32_t drop_thr, float drop_scale, uint64_t seed,
                                               uint32_t utt, uint32_t ph, uint32_t step, uint32_t layer) {
#pragma unroll
  for (int h8 = 0; h8 < 2; ++h8) {                         // 8 columns share one Philox call
    Philox4 rnd = Philox4{0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
    if (use_drop) rnd = dropout_words(seed, utt, ph, step, layer, (uint32_t)((col0 >> 3) + h8));
    const uint32_t wv[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
    const float4 ba = *reinterpret_cast<const float4*>(bias + col0 + 8 * h8);
    const float4 bb = *(reinterpret_cast<const float4*>(bias + col0 + 8 * h8) + 1);
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
#ifdef FCL_DEC_PROF
    if (dst)
#endif
    *reinterpret_cast<uint4*>(dst + ((size_t)((col0 >> 3) + h8) * 128 + r) * 16) = w;
  }
}

__global__ void __launch_bounds__(kDbThreads, 1)
decoder_bf16_pair_v1_kernel(FclDecoderBf16ParamsEx p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ DbShared sh;
#ifdef FCL_DEC_PROF
  const int role_rot = DB_WHATIF(p, 20) ? 128 : 0;        // what-if: the single-thread roles in the LAST four warps
#else
  constexpr int role_rot = 0;
#endif
  // logical thread id: warps 0-3 = single-thread roles, 4-19 = epilogue (physical warp % 4 == logical warp % 4)
  const int tid = ((int)threadIdx.x + role_rot) % kDbThreads, warp = tid >> 5, lane = tid & 31;
  const int H = p.dunits, U = p.prenet_units, O = p.odim, E = p.eunits;
  DbDims dm;
  dm.kU = U / 64; dm.kH = H / 64; dm.kE = E / 64; dm.gate_chunks = 4 * H / 256;
  dm.C = 1; dm.cr = 0;                                  // pair mode: every CTA owns all chunks of ITS tile
  const uint32_t rank = cluster_ctarank();              // 0 = leader (issues the MMAs), 1 = peer
  const int grp = (int)blockIdx.x >> 1;                 // the pair = one schedule slot; it walks SUPER-tiles (2 tiles)
  uint8_t* act = reinterpret_cast<uint8_t*>(p.act_priv) + (size_t)blockIdx.x * db_priv_bytes(U);          // x1 | x2
  uint8_t* zsh = reinterpret_cast<uint8_t*>(p.act_shared) + (size_t)blockIdx.x * db_shared_bytes(H);      // z images
  float* cws = p.c_ws + (size_t)blockIdx.x * 2 * H * 128;

  // this CTA's tile list (longest-processing-time schedule)
  if (tid == 0) sh.n_my_tiles = 0;
#ifdef FCL_DEC_PROF
  if (tid < 48) sh.prof[tid] = 0u;
#endif
  __syncthreads();
  const int n_super = (p.n_tiles + 1) >> 1;
  for (int t = tid; t < n_super; t += kDbThreads) {
    if (p.tile_slot[t] == grp) {
      const int k = p.tile_rank[t];
      if (k < kDbMaxTilesPerCta) { sh.my_tiles[k] = t; atomicMax(&sh.n_my_tiles, k + 1); }
    }
  }
  // Epilogue constants (gate biases, position column, prenet biases) staged in shared memory behind the ring when they
  // fit (S: 14 KB): the epilogue reads two float4 per hidden unit and chunk, and with ~200 KB of the SM's unified
  // L1 / shared memory given to the ring they kept missing the few KB of L1 that are left (ncu source view: the
  // FFMAs consuming them carried 16 % of all stall samples). Larger models read them from global memory as before.
  const float* b0s = p.b0, *wposs = p.wpos, *b1s = p.b1, *bp0s = p.bp0, *bp1s = p.bp1;
  if (p.smem_consts) {
    float* sc = reinterpret_cast<float*>(smem + (size_t)kDbStages * kStageBytes);
    for (int i = tid; i < 4 * H; i += kDbThreads) { sc[i] = p.b0[i]; sc[4 * H + i] = p.wpos[i]; sc[8 * H + i] = p.b1[i]; }
    for (int i = tid; i < U; i += kDbThreads) { sc[12 * H + i] = p.bp0[i]; sc[12 * H + U + i] = p.bp1[i]; }
    b0s = sc; wposs = sc + 4 * H; b1s = sc + 8 * H; bp0s = sc + 12 * H; bp1s = sc + 12 * H + U;
  }
  if (tid == 0) {
    for (int s = 0; s < kDbStages; ++s) { mbar_init(&sh.full[s], 2); mbar_init(&sh.empty[s], 1); }   // full: A + B producer
    for (int b = 0; b < 2; ++b) { mbar_init(&sh.tmem_full[b], 1); mbar_init(&sh.tmem_empty[b], kEpiThreads / 32); }
    for (int i = 0; i < 4; ++i) mbar_init(&sh.a_ready[i], kEpiThreads / 32);
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
#ifdef FCL_DEC_PROF
  const int n_ring = ((p.inflight >> 12) & 7) ? ((p.inflight >> 12) & 7) : kDbStages;   // what-if: shallower ring
#else
  constexpr int n_ring = kDbStages;
#endif
  // per-CTA halves: 128 of the 256 gate/prenet columns, 64 of the feat_out columns (odim zero-padded to 128)
  const uint32_t b_bytes_wide = 128u * 64u * 2u, b_bytes_feat = 64u * 64u * 2u;

  // Register budget (see decoder_bf16.cu): warps 0-3 give up 32 registers per thread, the epilogue threads get 16 more.
  // (104 is the most the pool gives: 112 compiles with 16 B of spills instead of 28 B but setmaxnreg.inc never returns.)
  // Each setmaxnreg has to dominate the code it is meant for, hence the two-level role dispatch.
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
  if (warp == 0) {
    // ================================================================ producer of the A operands (activation images)
    // Two producer threads share the ring (round 2): this one copies the A stages and waits for the operands the
    // epilogue produces; warp 2 copies the weight halves, which never depend on anything and so run ahead as far as the
    // ring allows. Each arrives once per slot with its own byte count (the `full` barriers count 2). One thread doing
    // both was the bottleneck of the operand stream: tools/decoder_prof.py showed it BUSY ~600 cycles per item (two
    // UBLKCP, expect_tx, the waits) next to the 16 epilogue warps, against 512 tensor cycles per item, with the ring
    // rarely full (profiles/r02_decoder_whatif.md).
    if (elect_one()) {
      uint32_t stage = 0, sphase = 0;                 // ring position / parity
      uint32_t rdy[4] = {0, 0, 0, 0};                 // parity of each a_ready barrier
      const bool prof = DB_PROF_ON(p);
      uint32_t pt = clk32();
      const uint32_t a_bytes = DB_WHATIF(p, 8) ? 0u : kABytes;
      // one operand image (segment) = n consecutive K stages: a tight loop per segment keeps this thread's
      // per-item instruction chain short (a single thread retires ~1 instruction per 10 cycles)
      auto seg = [&](const uint8_t* src, int n) {
        for (int ks = 0; ks < n; ++ks) {
          mbar_wait(&sh.empty[stage], sphase ^ 1u);
          mbar_arrive_expect_tx(&sh.full[stage], a_bytes);
          if (a_bytes) bulk_g2s(smem + (size_t)stage * kStageBytes, src, kABytes, &sh.full[stage]);
          src += kABytes;
          if (++stage == (uint32_t)n_ring) { stage = 0; sphase ^= 1u; }
        }
      };
      auto wait_operand = [&](int phase) {            // operand image written by the previous phase's epilogue
        DB_PROF(47);
        mbar_wait(&sh.a_ready[phase], rdy[phase]);
        rdy[phase] ^= 1u;
        DB_PROF(24 + phase);
      };
      for (int tk = 0, st; (st = DB_TILE(tk)) >= 0; ++tk) {
        const int steps = min(max(p.dur[p.order[(size_t)st * 256]], 0), FCL_MAX_DURATION);   // the pair runs the longer tile's steps
        const int tile = min(2 * st + (int)rank, p.n_tiles - 1);    // an odd tile count leaves the last peer a dummy (masked) tile
        const uint8_t* himg = reinterpret_cast<const uint8_t*>(p.hn_img) + (size_t)tile * E * 128 * 2;
        for (int m = 0; m < steps; ++m) {
          const int zp = m & 1;
          const uint8_t* z0cur = zsh + db_z_off(H, 0, zp), *z0new = zsh + db_z_off(H, 0, zp ^ 1);
          const uint8_t* z1cur = zsh + db_z_off(H, 0, 2 + zp), *z1new = zsh + db_z_off(H, 0, 2 + (zp ^ 1));
          const bool last_step = m + 1 == steps || p.tf_x1 != nullptr;
          // P1: [x1]
          if (tid == 0) db_trace(p, 600);
          wait_operand(0);
          seg(act + db_x1_off(), dm.kU);
          DB_PROF(28);
          // L0: [h | z0 | x2]
          if (tid == 0) db_trace(p, 601);
          for (int c = 0; c < dm.gate_chunks; ++c) {
            seg(himg, dm.kE);
            seg(z0cur, dm.kH);
            if (c == 0) { DB_PROF(29); wait_operand(1); }
            seg(act + db_x2_off(U), dm.kU);
          }
          DB_PROF(29);
          // L1: [z1 | z0']
          if (tid == 0) db_trace(p, 602);
          for (int c = 0; c < dm.gate_chunks; ++c) {
            seg(z1cur, dm.kH);
            if (c == 0) { DB_PROF(30); wait_operand(2); }
            seg(z0new, dm.kH);
          }
          DB_PROF(30);
          // FP: [h | z1'], feat_out and (unless it is the tile's last step / teacher forcing) the composed prenet.0
          if (tid == 0) db_trace(p, 603);
          for (int c = 0; c < (last_step ? 1 : 2); ++c) {
            seg(himg, dm.kE);
            if (c == 0) { DB_PROF(31); wait_operand(3); }
            seg(z1new, dm.kH);
          }
          DB_PROF(31);
        }
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ================================================================ producer of the B operands (this CTA's weight halves)
    if (elect_one()) {
      uint32_t stage = 0, sphase = 0;
      for (int tk = 0, st; (st = DB_TILE(tk)) >= 0; ++tk) {
        const int steps = min(max(p.dur[p.order[(size_t)st * 256]], 0), FCL_MAX_DURATION);
        for (int m = 0; m < steps; ++m) {
          for (int phase = 0; phase < 4; ++phase) {
            const int nch = dm.nchunks(phase, m + 1 == steps || p.tf_x1 != nullptr), kst = dm.kstages(phase);
            for (int c = 0; c < nch; ++c) {
              const uint32_t bb = (phase == 3 && c == 0) ? b_bytes_feat : b_bytes_wide;
              // stage blocks hold both halves back to back: [rank 0 half][rank 1 half]
              const uint8_t* wptr = reinterpret_cast<const uint8_t*>(p.w_stream) +
                                    dm.w_off(phase, c, 2 * b_bytes_wide, 2 * b_bytes_feat) + (size_t)rank * bb;
              for (int ks = 0; ks < kst; ++ks) {
                mbar_wait(&sh.empty[stage], sphase ^ 1u);
                mbar_arrive_expect_tx(&sh.full[stage], bb);
                bulk_g2s(smem + (size_t)stage * kStageBytes + kABytes, wptr, bb, &sh.full[stage]);
                wptr += 2 * bb;
                if (++stage == n_ring) { stage = 0; sphase ^= 1u; }
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
      const bool prof = DB_PROF_ON(p);
      uint32_t pt = clk32();
      const uint32_t pt_start = pt; (void)pt_start;
      const uint32_t idesc_wide = idesc_op_f32(256u, 256u), idesc_feat = idesc_op_f32(256u, 128u);
      // descriptors built incrementally (see decoder_bf16.cu): low word = (address >> 4) | (LBO >> 4) << 16
      const uint32_t ring_lo = smem_u32(smem) >> 4;
      constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);
      constexpr uint32_t kALo = (2048u >> 4) << 16;
      uint32_t s_lo = ring_lo;
      for (int tk = 0, st; (st = DB_TILE(tk)) >= 0; ++tk) {
        const int steps = min(max(p.dur[p.order[(size_t)st * 256]], 0), FCL_MAX_DURATION);
        for (int m = 0; m < steps; ++m) {
          for (int phase = 0; phase < 4; ++phase) {
            const int nch = dm.nchunks(phase, m + 1 == steps || p.tf_x1 != nullptr), kst = dm.kstages(phase);
            for (int c = 0; c < nch; ++c) {
              const bool feat = phase == 3 && c == 0;
              const uint32_t idesc = feat ? idesc_feat : idesc_wide;
              const uint32_t b_lbo = feat ? 64u * 16u : 128u * 16u;          // rows of THIS CTA's half x 16 B
              const uint32_t b_lo = (kABytes >> 4) + ((b_lbo >> 4) << 16), b_kstep = (2u * b_lbo) >> 4;
              const uint32_t buf = chunk_ctr & 1u, use = chunk_ctr >> 1;
              if (rank == 0) {
                mbar_wait(&sh.tmem_empty[buf], (use & 1u) ^ 1u);
                if (use > 0) mbar_wait(&sh.peer_tmem_empty[buf], (use & 1u) ^ 1u);   // completion #(use-1): the peer drained it too
                tc_fence_after();
                db_trace(p, 100 + phase * 10 + c);
                DB_PROF(4 * phase);
              }
              const uint32_t d_tmem = tmem + buf * 256u;
              for (int ks = 0; ks < kst; ++ks) {
                mbar_wait(&sh.full[stage], sphase);
                if (rank == 1) {
                  mbar_arrive_remote(&sh.peer_full[stage], 0);               // tell the leader our half has landed
                } else {
                  mbar_wait(&sh.peer_full[stage], sphase);
                  tc_fence_after();
                  if (ks == 0) db_trace(p, 200 + phase * 10 + c);
                  DB_PROF(4 * phase + (ks == 0 ? 1 : 2));
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    const uint64_t ad = ((uint64_t)kDescHi << 32) | (s_lo + (uint32_t)k * (4096u >> 4) + kALo);
                    const uint64_t bd = ((uint64_t)kDescHi << 32) | (s_lo + b_lo + (uint32_t)k * b_kstep);
                    mma2_bf16_ss(d_tmem, ad, bd, idesc, (ks > 0 || k > 0) ? 1u : 0u);
                  }
                  mma2_commit(&sh.empty[stage]);                             // frees the stage in BOTH CTAs
                  DB_PROF(4 * phase + 3);
#ifdef FCL_DEC_PROF
                  if (prof) ++sh.prof[34 + phase];
#endif
                }
                s_lo += kStageBytes >> 4;
                if (++stage == n_ring) { stage = 0; sphase ^= 1u; s_lo = ring_lo; }
              }
              if (rank == 0) {
                mma2_commit(&sh.tmem_full[buf]);                             // accumulator ready in BOTH CTAs
                db_trace(p, 300 + phase * 10 + c);
              }
              ++chunk_ctr;
            }
          }
#ifdef FCL_DEC_PROF
          if (prof) ++sh.prof[32];
#endif
        }
      }
#ifdef FCL_DEC_PROF
      if (prof) sh.prof[33] = clk32() - pt_start;
#endif
    }
    __syncwarp();
  } else if (warp == 3) {
    // ================================================================ peer only: "accumulator drained" relay
    if (rank == 1 && elect_one()) {
      uint32_t chunk_ctr = 0;
      for (int tk = 0, st; (st = DB_TILE(tk)) >= 0; ++tk) {
        const int steps = min(max(p.dur[p.order[(size_t)st * 256]], 0), FCL_MAX_DURATION);
        for (int m = 0; m < steps; ++m) {
          for (int phase = 0; phase < 4; ++phase) {
            const int nch = dm.nchunks(phase, m + 1 == steps || p.tf_x1 != nullptr);
            for (int c = 0; c < nch; ++c) {
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
    const bool use_drop = p.dropout_p > 0.f && !DB_WHATIF(p, 11);
    const bool wi_noc = DB_WHATIF(p, 9), wi_noimg = DB_WHATIF(p, 10);
    const bool wi_skel = DB_WHATIF(p, 15), wi_nomath = DB_WHATIF(p, 16); (void)wi_nomath;
    const uint32_t drop_thr = dropout_threshold16(p.dropout_p);
    const float drop_scale = use_drop ? 1.0f / (1.0f - p.dropout_p) : 1.0f;
    const bool prof = DB_PROF_ON(p) && tid == 128;
    uint32_t pt = clk32();

    for (int tk = 0, st; (st = DB_TILE(tk)) >= 0; ++tk) {
      const int tile = 2 * st + (int)rank;
      const int sidx = tile * 128 + r;
      int row = -1, d = 0, foff = 0, utt = 0, ph = 0;
      if (tile < p.n_tiles && sidx < p.n_rows) {
        row = p.order[sidx];
        d = min(max(p.dur[row], 0), FCL_MAX_DURATION);
        foff = p.frame_off[row];
        utt = p.row_utt[row];
        ph = p.row_phone[row];
      }
      const int steps = min(max(p.dur[p.order[(size_t)st * 256]], 0), FCL_MAX_DURATION);
      if (steps == 0) continue;
      const int zset = dm.C > 1 ? (tk & 1) : 0;      // one set is enough without a group (keeps the scratch L2-resident)
      // ---- tile init: x1 of step 0 (the first input frame is zero: prenet.0 sees only its bias) and zero z images
      {
#pragma unroll 1
        for (int g = 0; g < 4; ++g)
          prenet_store16(nullptr, bp0s, cs * 64 + g * 16, r, act + db_x1_off(), use_drop, drop_thr, drop_scale,
                         p.dropout_seed, (uint32_t)utt, (uint32_t)ph, 0u, 0u);
        const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
        // (in group mode every CTA of the group zeroes the same images with the same zeros; the set alternates per
        // tile so a CTA that is one tile ahead never touches images a slower CTA still reads)
        for (int kc = cs; kc < H / 8; kc += 4) {
          *reinterpret_cast<uint4*>(zsh + db_z_off(H, zset, 0) + ((size_t)kc * 128 + r) * 16) = z4;
          *reinterpret_cast<uint4*>(zsh + db_z_off(H, zset, 2) + ((size_t)kc * 128 + r) * 16) = z4;
        }
        fence_proxy_async_global();
        warp_arrive(&sh.a_ready[0], lane);
      }

      for (int m = 0; m < steps; ++m) {
        const int zp = m & 1;
        uint8_t* z0cur = zsh + db_z_off(H, zset, zp), *z0new = zsh + db_z_off(H, zset, zp ^ 1);
        uint8_t* z1cur = zsh + db_z_off(H, zset, 2 + zp), *z1new = zsh + db_z_off(H, zset, 2 + (zp ^ 1));
        const float pos = (row >= 0 && m < d) ? __fdiv_rn((float)m, (float)d) : 0.f;

        // ---------------- P1: prenet layer 1 (bias, ReLU, dropout) -> x2 image; 64 columns per thread
        {
          const uint32_t buf = chunk_ctr & 1u, use = chunk_ctr >> 1;
          DB_PROF(46);
          mbar_wait(&sh.tmem_full[buf], use & 1u);
          tc_fence_after();
          DB_PROF(16);
          db_dummy_work(p, (uint32_t)tid + chunk_ctr);
          if (tid == 128) db_trace(p, 400);
#pragma unroll 1
          for (int g = 0; g < (wi_skel ? 0 : 4); ++g) {
            float v[16];
            const int col0 = cs * 64 + g * 16;
            tmem_ld16(lane_addr + buf * 256u + (uint32_t)col0, v);
            prenet_store16(v, bp1s, col0, r, wi_noimg ? nullptr : act + db_x2_off(U), use_drop, drop_thr, drop_scale, p.dropout_seed,
                           (uint32_t)utt, (uint32_t)ph, (uint32_t)m, 1u);
          }
          tc_fence_before();
          warp_arrive(&sh.tmem_empty[buf], lane);
          ++chunk_ctr;
          fence_proxy_async_global();
          warp_arrive(&sh.a_ready[1], lane);
          DB_PROF(17);
          if (tid == 128) db_trace(p, 500);
        }

        // ---------------- L0, L1: zoneout LSTM cells; per chunk this thread owns 16 hidden units of its row.
#pragma unroll 1
        for (int layer = 0; layer < 2; ++layer) {
          const uint8_t* zcur = layer == 0 ? z0cur : z1cur;
          uint8_t* znew = layer == 0 ? z0new : z1new;
          float* cl = cws + (size_t)layer * H * 128;
          const float* bias = layer == 0 ? b0s : b1s;
          float c_cur[16];
          uint4 z_cur[2];
#pragma unroll 1
          for (int c = 0; c < dm.gate_chunks; ++c) {
            const int u0 = c * 64 + cs * 16;                          // first of this thread's 16 hidden units
            // old cell state / old z of this chunk: requested BEFORE waiting for the accumulator (their L2 latency hides
            // behind the MMAs when the tensor pipe is the longer pole). No second register set for the next chunk: at 96
            // registers per thread it is spilled (tried three times: a second set; re-loading each 4-unit group's registers
            // for the next chunk as soon as the group is done -- 280 B of spills, 1.80 -> 2.19 ms; the same 2 units at a
            // time with 8-column TMEM loads -- 64 B of spills, 1.73 -> 1.76 ms).
#pragma unroll
            for (int j = 0; j < 16; ++j) c_cur[j] = (m == 0 || wi_noc || wi_skel) ? 0.f : __ldcg(cl + (size_t)(u0 + j) * 128 + r);
            if (!wi_skel) {
              z_cur[0] = __ldcg(reinterpret_cast<const uint4*>(zcur + ((size_t)(u0 >> 3) * 128 + r) * 16));
              z_cur[1] = __ldcg(reinterpret_cast<const uint4*>(zcur + ((size_t)((u0 >> 3) + 1) * 128 + r) * 16));
            } else {
              z_cur[0] = z_cur[1] = make_uint4(0u, 0u, 0u, 0u);
            }
            const uint32_t buf = chunk_ctr & 1u, use = chunk_ctr >> 1;
            DB_PROF(19 + 2 * layer);
            mbar_wait(&sh.tmem_full[buf], use & 1u);
            tc_fence_after();
            DB_PROF(18 + 2 * layer);
            db_dummy_work(p, (uint32_t)tid + chunk_ctr);
            if (tid == 128) db_trace(p, 400 + (1 + layer) * 10 + c);
            uint32_t zout[8];
#ifdef FCL_DEC_PROF
            for (int g = 0; g < 8; ++g) zout[g] = 0u;
            if (!wi_skel)
#endif
#pragma unroll
            for (int g = 0; g < 4; ++g) {                             // 4 units (16 accumulator columns) at a time
              float v[16];
              tmem_ld16(lane_addr + buf * 256u + (uint32_t)(cs * 64 + g * 16), v);
              float zn[4];
#ifdef FCL_DEC_PROF
              if (wi_nomath) {
                for (int j = 0; j < 4; ++j) zn[j] = v[4 * j] + v[4 * j + 1] + v[4 * j + 2] + v[4 * j + 3];
              } else
#endif
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int ul = g * 4 + j, u = u0 + ul;
                float4 add = *reinterpret_cast<const float4*>(bias + 4 * u);
                if (layer == 0) {
                  const float4 wp = *reinterpret_cast<const float4*>(wposs + 4 * u);
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
                c_cur[ul] = fmaf(zo, cold, zk * cn);                  // new cell state: stored after the hand-overs below
              }
              zout[2 * g] = pack_op(zn[0], zn[1]);
              zout[2 * g + 1] = pack_op(zn[2], zn[3]);
            }
            tc_fence_before();
            warp_arrive(&sh.tmem_empty[buf], lane);
            ++chunk_ctr;
            if (!wi_noimg && !wi_skel) {
              *reinterpret_cast<uint4*>(znew + ((size_t)(u0 >> 3) * 128 + r) * 16) = make_uint4(zout[0], zout[1], zout[2], zout[3]);
              *reinterpret_cast<uint4*>(znew + ((size_t)((u0 >> 3) + 1) * 128 + r) * 16) = make_uint4(zout[4], zout[5], zout[6], zout[7]);
            }
            if (c + 1 == dm.gate_chunks) {                            // the next phase's operand is complete: signal it BEFORE
              fence_proxy_async_global();                             // the cell-state stores join the write queue the fence drains
              warp_arrive(&sh.a_ready[2 + layer], lane);
            }
            if (!wi_noc && !wi_skel) {
#pragma unroll
              for (int j = 0; j < 16; ++j) cl[(size_t)(u0 + j) * 128 + r] = c_cur[j];
            }
            if (tid == 128) db_trace(p, 500 + (1 + layer) * 10 + c);
          }
          DB_PROF(19 + 2 * layer);
        }

        // ---------------- FP chunk 0: feat_out -> output frame, stored straight to its final (ragged) position
        if (dm.owns(3, 0)) {
          const uint32_t buf = chunk_ctr & 1u, use = chunk_ctr >> 1;
          mbar_wait(&sh.tmem_full[buf], use & 1u);
          tc_fence_after();
          DB_PROF(22);
          db_dummy_work(p, (uint32_t)tid + chunk_ctr);
          if (tid == 128) db_trace(p, 430);
          for (int g = cs; g < (wi_skel ? 0 : O / 16); g += 4) {     // 16-column groups dealt over the 4 column sets
            float v[16];
            tmem_ld16(lane_addr + buf * 256u + (uint32_t)(g * 16), v);
            if (row >= 0 && m < d) {                                 // exhausted rows are masked (decoder_sa.py:625-629)
              float4* o = reinterpret_cast<float4*>(p.before + ((size_t)foff + m) * O + g * 16);
#pragma unroll
              for (int qd = 0; qd < 4; ++qd)                          // streaming store: written once, read by the next kernel
                __stcs(o + qd, make_float4(v[4 * qd], v[4 * qd + 1], v[4 * qd + 2], v[4 * qd + 3]));
            }
          }
          tc_fence_before();
          warp_arrive(&sh.tmem_empty[buf], lane);
          ++chunk_ctr;
          DB_PROF(23);
          if (tid == 128) db_trace(p, 530);
        }
        // ---------------- teacher forcing: the next step's x1 image comes from the ground-truth frame (fcl_prenet0_tf)
        if (p.tf_x1 && m + 1 < steps) {
          const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.tf_x1) + ((size_t)foff + m) * U + cs * 64);
          const bool live = row >= 0 && m + 1 < d;
#pragma unroll
          for (int k8 = 0; k8 < 8; ++k8) {
            const uint4 w = live ? __ldg(src + k8) : make_uint4(0u, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(act + db_x1_off() + ((size_t)(cs * 8 + k8) * 128 + r) * 16) = w;
          }
          fence_proxy_async_global();
          warp_arrive(&sh.a_ready[0], lane);
        }
        // ---------------- FP chunk 1: prenet layer 0 of the NEXT step (composed with feat_out) -> x1 image
        if (m + 1 < steps && !p.tf_x1) {
          const uint32_t buf = chunk_ctr & 1u, use = chunk_ctr >> 1;
          mbar_wait(&sh.tmem_full[buf], use & 1u);
          tc_fence_after();
          DB_PROF(44);
          db_dummy_work(p, (uint32_t)tid + chunk_ctr);
          if (tid == 128) db_trace(p, 431);
#pragma unroll 1
          for (int g = 0; g < (wi_skel ? 0 : 4); ++g) {
            float v[16];
            const int col0 = cs * 64 + g * 16;
            tmem_ld16(lane_addr + buf * 256u + (uint32_t)col0, v);
            prenet_store16(v, bp0s, col0, r, wi_noimg ? nullptr : act + db_x1_off(), use_drop, drop_thr, drop_scale, p.dropout_seed,
                           (uint32_t)utt, (uint32_t)ph, (uint32_t)(m + 1), 0u);
          }
          tc_fence_before();
          warp_arrive(&sh.tmem_empty[buf], lane);
          ++chunk_ctr;
          fence_proxy_async_global();
          warp_arrive(&sh.a_ready[0], lane);
          DB_PROF(45);
          if (tid == 128) db_trace(p, 531);
        }
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();                                   // the leader's MMAs read the peer's shared memory: leave together
  if (warp == 2) tmem_dealloc2(tmem, 512);
#ifdef FCL_DEC_PROF
  if (DB_PROF_ON(p) && tid < 48) p.trace[tid] = sh.prof[tid];
#endif
}

}  // namespace pair_v1


}  // namespace fcl

extern "C" int fcl_decoder_bf16_pair_v1(const FclDecoderBf16Params* p, void* stream) {
  using namespace fcl;
  using namespace fcl::pair_v1;
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
#ifdef FCL_DEC_PROF
  if ((p->inflight >> 22) & 1) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)80 << 20);   // what-if: evict_last set-aside
  if ((p->inflight >> 23) & 1) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
#endif
  size_t smem = (size_t)kDbStages * kStageBytes;
  const size_t consts = (size_t)(12 * p->dunits + 2 * p->prenet_units) * sizeof(float);
  FclDecoderBf16ParamsEx px;
  static_cast<FclDecoderBf16Params&>(px) = *p;
  px.smem_consts = consts <= 32 * 1024 ? 1 : 0;
#ifdef FCL_DEC_PROF
  if ((p->inflight >> 24) & 1) px.smem_consts = 0;        // what-if: constants from global memory as before
#endif
  if (px.smem_consts) smem += consts;
  if (int rc = ensure_dyn_smem(decoder_bf16_pair_v1_kernel, smem, "fcl_decoder_bf16_pair_v1")) return rc;
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
  cudaError_t e = cudaLaunchKernelEx(&cfg, decoder_bf16_pair_v1_kernel, px);
  if (e != cudaSuccess) { set_error("fcl_decoder_bf16_pair_v1: %s", cudaGetErrorString(e)); return FCL_ECUDA; }
  return check_launch("fcl_decoder_bf16_pair_v1");
}
