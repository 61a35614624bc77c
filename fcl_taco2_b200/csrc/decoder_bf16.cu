// K4 (tensor-core path): persistent decoder step loop on tcgen05.
// Reference: nets/modules/decoder_sa.py:577-617 (loop), :146-158 (Prenet, always-on dropout),
// :63-96 (ZoneOutCell around torch.nn.LSTMCell), :398 (feat_out), fused with the ragged gather :619-630.
//
// One CTA owns a tile of 128 duration-sorted phoneme rows for all of their steps (rows are independent, so
// no inter-CTA communication exists). Per step it runs four dependent GEMM phases on the tensor cores
//   P1 prenet.1: x1 (K = 256) x 256
//   L0 gates of cell 0: [h | z0 | x2] (K = E + H + 256) x 4H      L1 gates of cell 1: [z1 | z0'] (K = 2H) x 4H
//   FP [h | z1'] (K = E + H) x {80 feat_out columns ; 256 columns of prenet.0 composed with feat_out}
// Two algebraic rearrangements keep the dependency chain short: (1) feat_out has no activation, so
// prenet.0(y) = relu(y Wp0^T + b) = relu([z1'|h] (Wfeat^T Wp0^T) + b): the next step's prenet.0 shares the
// operand (and the phase) of feat_out; (2) inside every phase the K-slices that are already known (h, the
// previous step's z) come first and the slice produced by the previous phase comes last, so the MMAs of a
// phase start while the previous phase's epilogue is still running.
// All GEMMs use bf16 operands, fp32 accumulators in TMEM (two 256-column buffers: the epilogue of chunk j overlaps
// the MMAs of chunk j+1), fp32 cell state. The encoder state h of the tile is a K-slice of the A operand
// (it is NOT hoisted into a per-row fp32 table: re-reading such a table every step made the epilogue
// latency-bound on HBM). Weights (bf16, pre-tiled as UMMA core matrices, in consumption order) and the
// activation operands stream through a 4-stage shared-memory ring with cp.async.bulk; the activation
// operands of a tile (x0,x1,x2,z0,z1 as bf16 core-matrix images) live in a per-CTA global scratch that stays
// L2-resident, written by the epilogue warps and re-read by the bulk copies (generic->async proxy fence +
// mbarrier hand-over).
//
// Warp roles (640 threads): warp 0 = bulk-copy producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
// warps 4-19 = epilogue (TMEM lane quarter = warp % 4; column quarter = (warp - 4) / 4: 64 of a chunk's 256
// columns). Gate columns are interleaved (unit*4 + {i,f,g,o}) so one thread owns whole LSTM cells; the old
// cell state / old z of a chunk are requested right before the thread waits for that chunk's accumulator.
// setmaxnreg gives the epilogue threads 104 registers and leaves 64 to warps 0-3 (see the role dispatch).
#include "common.cuh"
#include "umma.cuh"

namespace fcl {
using namespace umma;

constexpr int kDbThreads = 640;
constexpr int kDbStages = 4;
constexpr uint32_t kABytes = 128u * 64u * 2u;           // one A stage: 128 rows x 64 k (bf16)
constexpr uint32_t kBBytesMax = 256u * 64u * 2u;        // one B stage: up to 256 cols x 64 k
constexpr uint32_t kStageBytes = kABytes + kBBytesMax;  // 48 KB
constexpr int kEpiThreads = 512;
constexpr int kDbMaxTilesPerCta = 512;

struct DbShared {
  uint64_t full[kDbStages], empty[kDbStages];
  uint64_t tmem_full[2], tmem_empty[2];
  uint64_t a_ready[4];        // x1, x2, z0', z1' operand images complete (epilogue -> producer)
  uint32_t tmem_base;
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
  if (p.trace && blockIdx.x == 0) {
    const unsigned long long n = atomicAdd(reinterpret_cast<unsigned long long*>(p.trace), 1ull);
    if (n < (unsigned long long)p.trace_cap) {
      p.trace[2 + 2 * n] = id;
      p.trace[3 + 2 * n] = clock64();
    }
  }
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
decoder_bf16_kernel(FclDecoderBf16Params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ DbShared sh;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = p.dunits, U = p.prenet_units, O = p.odim, E = p.eunits;
  DbDims dm;
  dm.kU = U / 64; dm.kH = H / 64; dm.kE = E / 64; dm.gate_chunks = 4 * H / 256;
  dm.C = p.group; dm.cr = (int)blockIdx.x % p.group;
  const int grp = (int)blockIdx.x / p.group;
  uint8_t* act = reinterpret_cast<uint8_t*>(p.act_priv) + (size_t)blockIdx.x * db_priv_bytes(U);          // x1 | x2
  uint8_t* zsh = reinterpret_cast<uint8_t*>(p.act_shared) + (size_t)grp * db_shared_bytes(H);             // z images
  float* cws = p.c_ws + (size_t)blockIdx.x * 2 * H * 128;

  // this CTA's tile list (longest-processing-time schedule)
  if (tid == 0) sh.n_my_tiles = 0;
  __syncthreads();
  for (int t = tid; t < p.n_tiles; t += kDbThreads) {
    if (p.tile_slot[t] == grp) {
      const int k = p.tile_rank[t];
      if (k < kDbMaxTilesPerCta) { sh.my_tiles[k] = t; atomicMax(&sh.n_my_tiles, k + 1); }
    }
  }
  if (tid == 0) {
    for (int s = 0; s < kDbStages; ++s) { mbar_init(&sh.full[s], 1); mbar_init(&sh.empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&sh.tmem_full[b], 1); mbar_init(&sh.tmem_empty[b], kEpiThreads / 32); }
    for (int i = 0; i < 4; ++i) mbar_init(&sh.a_ready[i], kEpiThreads / 32);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&sh.tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh.tmem_base;
  const uint32_t b_bytes_wide = 256u * 64u * 2u, b_bytes_feat = (uint32_t)O * 64u * 2u;

  // Register budget: the launch gives every thread 96 (65536 / 640, rounded down); the single-thread roles of warps
  // 0-3 need far fewer, the epilogue (16 LSTM cells per thread in flight) spills at 96. Warps 0-3 release 32 each
  // (4096 in total), the 512 epilogue threads take 8 more each (4096). Asking for 112 (exactly the 8192 that are free
  // on paper) never returns: setmaxnreg.inc blocks until the pool can serve it, so keep a margin.
  // (each setmaxnreg has to dominate the code it is meant for, hence the two-level role dispatch)
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
  if (warp == 0) {
    // ================================================================ producer
    if (elect_one()) {
      uint32_t stage = 0, sphase = 0;                 // ring position / parity
      uint32_t rdy[4] = {0, 0, 0, 0};                 // parity of each a_ready barrier
      int sync_ev[2] = {0, 0};                        // group barriers passed so far (z0', z1')
      for (int tk = 0, tile; (tile = DB_TILE(tk)) >= 0; ++tk) {
        const int steps = min(max(p.dur[p.order[(size_t)tile * 128]], 0), FCL_MAX_DURATION);
        const uint8_t* himg = reinterpret_cast<const uint8_t*>(p.hn_img) + (size_t)tile * E * 128 * 2;
        const int zset = dm.C > 1 ? (tk & 1) : 0;      // one set is enough without a group (keeps the scratch L2-resident)
        for (int m = 0; m < steps; ++m) {
          const int zp = m & 1;
          const uint8_t* z0cur = zsh + db_z_off(H, zset, zp), *z0new = zsh + db_z_off(H, zset, zp ^ 1);
          const uint8_t* z1cur = zsh + db_z_off(H, zset, 2 + zp), *z1new = zsh + db_z_off(H, zset, 2 + (zp ^ 1));
          for (int phase = 0; phase < 4; ++phase) {
            if (tid == 0) db_trace(p, 600 + phase);
            const int nch = dm.nchunks(phase, m + 1 == steps || p.tf_x1 != nullptr), kst = dm.kstages(phase), late = dm.late_stage(phase);
            bool waited = false;
            if (dm.C > 1 && phase >= 2) {
              // group mode: tell the other CTAs of the tile that this CTA's slice of z0' (z1') is written
              mbar_wait(&sh.a_ready[phase], rdy[phase]);
              rdy[phase] ^= 1u;
              __threadfence();
              atomicAdd(p.group_sync + 2 * grp + (phase - 2), 1);
              ++sync_ev[phase - 2];
            }
            for (int c = 0; c < nch; ++c) {
              if (!dm.owns(phase, c)) continue;
              const uint32_t bb = (phase == 3 && c == 0) ? b_bytes_feat : b_bytes_wide;
              const uint8_t* wptr = reinterpret_cast<const uint8_t*>(p.w_stream) + dm.w_off(phase, c, b_bytes_wide, b_bytes_feat);
              for (int ks = 0; ks < kst; ++ks) {
                if (!waited && ks == late) {           // operand written by the previous phase's epilogue(s)
                  waited = true;
                  if (dm.C > 1 && phase >= 2) {
                    const int target = sync_ev[phase - 2] * dm.C;
                    const volatile int* ctr = p.group_sync + 2 * grp + (phase - 2);
                    while (*ctr < target) __nanosleep(64);
                    __threadfence();
                    fence_proxy_async_all();
                  } else {
                    mbar_wait(&sh.a_ready[phase], rdy[phase]);
                    rdy[phase] ^= 1u;
                  }
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
                uint8_t* st = smem + (size_t)stage * kStageBytes;
                bulk_g2s(st, asrc, kABytes, &sh.full[stage]);
                bulk_g2s(st + kABytes, wptr, bb, &sh.full[stage]);
                wptr += bb;
                if (++stage == kDbStages) { stage = 0; sphase ^= 1u; }
              }
            }
            if (!waited && !(dm.C > 1 && phase >= 2)) {   // no owned chunk in this phase: keep the barrier parity in step
              mbar_wait(&sh.a_ready[phase], rdy[phase]);
              rdy[phase] ^= 1u;
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    if (elect_one()) {
      uint32_t stage = 0, sphase = 0;
      uint32_t chunk_ctr = 0;                          // accumulator buffer = chunk_ctr & 1
      const uint32_t idesc_wide = idesc_op_f32(128u, 256u), idesc_feat = idesc_op_f32(128u, (uint32_t)O);
      // Shared-memory descriptors, built incrementally: this thread shares its scheduler with four epilogue warps, so
      // every instruction between two tcgen05.mma counts. low word = (address >> 4) | (LBO >> 4) << 16 (addresses are
      // below 256 KB: no masking needed), high word = (SBO >> 4) | version 1 << 14, the same for every operand.
      const uint32_t ring_lo = smem_u32(smem) >> 4;
      constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);
      constexpr uint32_t kALo = (2048u >> 4) << 16;   // A: LBO = 128 rows x 16 B
      uint32_t s_lo = ring_lo;                         // descriptor address field of the current ring slot
      for (int tk = 0, tile; (tile = DB_TILE(tk)) >= 0; ++tk) {
        const int steps = min(max(p.dur[p.order[(size_t)tile * 128]], 0), FCL_MAX_DURATION);
        for (int m = 0; m < steps; ++m) {
          for (int phase = 0; phase < 4; ++phase) {
            const int nch = dm.nchunks(phase, m + 1 == steps || p.tf_x1 != nullptr), kst = dm.kstages(phase);
            for (int c = 0; c < nch; ++c) {
              if (!dm.owns(phase, c)) continue;
              const bool feat = phase == 3 && c == 0;
              const uint32_t ncols = feat ? (uint32_t)O : 256u;
              const uint32_t idesc = feat ? idesc_feat : idesc_wide;
              const uint32_t b_lbo = ncols * 16u;
              const uint32_t b_lo = (kABytes >> 4) + ((b_lbo >> 4) << 16), b_kstep = (2u * b_lbo) >> 4;
              const uint32_t buf = chunk_ctr & 1u, use = chunk_ctr >> 1;
              mbar_wait(&sh.tmem_empty[buf], (use & 1u) ^ 1u);
              tc_fence_after();
              db_trace(p, 100 + phase * 10 + c);
              const uint32_t d_tmem = tmem + buf * 256u;
              for (int ks = 0; ks < kst; ++ks) {
                mbar_wait(&sh.full[stage], sphase);
                tc_fence_after();
                if (ks == 0) db_trace(p, 200 + phase * 10 + c);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint64_t ad = ((uint64_t)kDescHi << 32) | (s_lo + (uint32_t)k * (4096u >> 4) + kALo);
                  const uint64_t bd = ((uint64_t)kDescHi << 32) | (s_lo + b_lo + (uint32_t)k * b_kstep);
                  mma_bf16_ss(d_tmem, ad, bd, idesc, (ks > 0 || k > 0) ? 1u : 0u);
                }
                mma_commit(&sh.empty[stage]);
                s_lo += kStageBytes >> 4;
                if (++stage == kDbStages) { stage = 0; sphase ^= 1u; s_lo = ring_lo; }
              }
              mma_commit(&sh.tmem_full[buf]);
              db_trace(p, 300 + phase * 10 + c);
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

    for (int tk = 0, tile; (tile = DB_TILE(tk)) >= 0; ++tk) {
      const int sidx = tile * 128 + r;
      int row = -1, d = 0, foff = 0, utt = 0, ph = 0;
      if (sidx < p.n_rows) {
        row = p.order[sidx];
        d = min(max(p.dur[row], 0), FCL_MAX_DURATION);
        foff = p.frame_off[row];
        utt = p.row_utt[row];
        ph = p.row_phone[row];
      }
      const int steps = min(max(p.dur[p.order[(size_t)tile * 128]], 0), FCL_MAX_DURATION);
      if (steps == 0) continue;
      const int zset = dm.C > 1 ? (tk & 1) : 0;      // one set is enough without a group (keeps the scratch L2-resident)
      // ---- tile init: x1 of step 0 (the first input frame is zero: prenet.0 sees only its bias) and zero z images
      {
#pragma unroll 1
        for (int g = 0; g < 4; ++g)
          prenet_store16(nullptr, p.bp0, cs * 64 + g * 16, r, act + db_x1_off(), use_drop, drop_thr, drop_scale,
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
          mbar_wait(&sh.tmem_full[buf], use & 1u);
          tc_fence_after();
          if (tid == 128) db_trace(p, 400);
#pragma unroll 1
          for (int g = 0; g < 4; ++g) {
            float v[16];
            const int col0 = cs * 64 + g * 16;
            tmem_ld16(lane_addr + buf * 256u + (uint32_t)col0, v);
            prenet_store16(v, p.bp1, col0, r, act + db_x2_off(U), use_drop, drop_thr, drop_scale, p.dropout_seed,
                           (uint32_t)utt, (uint32_t)ph, (uint32_t)m, 1u);
          }
          tc_fence_before();
          warp_arrive(&sh.tmem_empty[buf], lane);
          ++chunk_ctr;
          fence_proxy_async_global();
          warp_arrive(&sh.a_ready[1], lane);
          if (tid == 128) db_trace(p, 500);
        }

        // ---------------- L0, L1: zoneout LSTM cells; per chunk this thread owns 16 hidden units of its row.
#pragma unroll 1
        for (int layer = 0; layer < 2; ++layer) {
          const uint8_t* zcur = layer == 0 ? z0cur : z1cur;
          uint8_t* znew = layer == 0 ? z0new : z1new;
          float* cl = cws + (size_t)layer * H * 128;
          const float* bias = layer == 0 ? p.b0 : p.b1;
          float c_cur[16];
          uint4 z_cur[2];
#pragma unroll 1
          for (int c = dm.cr % dm.C; c < dm.gate_chunks; c += dm.C) {   // this CTA's chunks
            const int u0 = c * 64 + cs * 16;                          // first of this thread's 16 hidden units
            // old cell state / old z of this chunk: requested BEFORE waiting for the accumulator (their L2 latency hides
            // behind the MMAs). No second register set for the next chunk: at 96 registers per thread it was spilled
            // right after the loads, which made the "prefetch" a blocking load plus local-memory traffic.
#pragma unroll
            for (int j = 0; j < 16; ++j) c_cur[j] = m == 0 ? 0.f : __ldcg(cl + (size_t)(u0 + j) * 128 + r);
            z_cur[0] = __ldcg(reinterpret_cast<const uint4*>(zcur + ((size_t)(u0 >> 3) * 128 + r) * 16));
            z_cur[1] = __ldcg(reinterpret_cast<const uint4*>(zcur + ((size_t)((u0 >> 3) + 1) * 128 + r) * 16));
            const uint32_t buf = chunk_ctr & 1u, use = chunk_ctr >> 1;
            mbar_wait(&sh.tmem_full[buf], use & 1u);
            tc_fence_after();
            if (tid == 128) db_trace(p, 400 + (1 + layer) * 10 + c);
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
            if (tid == 128) db_trace(p, 500 + (1 + layer) * 10 + c);
          }
          fence_proxy_async_global();
          warp_arrive(&sh.a_ready[2 + layer], lane);
        }

        // ---------------- FP chunk 0: feat_out -> output frame, stored straight to its final (ragged) position
        if (dm.owns(3, 0)) {
          const uint32_t buf = chunk_ctr & 1u, use = chunk_ctr >> 1;
          mbar_wait(&sh.tmem_full[buf], use & 1u);
          tc_fence_after();
          if (tid == 128) db_trace(p, 430);
          for (int g = cs; g < O / 16; g += 4) {                     // 16-column groups dealt over the 4 column sets
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
          if (tid == 128) db_trace(p, 431);
#pragma unroll 1
          for (int g = 0; g < 4; ++g) {
            float v[16];
            const int col0 = cs * 64 + g * 16;
            tmem_ld16(lane_addr + buf * 256u + (uint32_t)col0, v);
            prenet_store16(v, p.bp0, col0, r, act + db_x1_off(), use_drop, drop_thr, drop_scale, p.dropout_seed,
                           (uint32_t)utt, (uint32_t)ph, (uint32_t)(m + 1), 0u);
          }
          tc_fence_before();
          warp_arrive(&sh.tmem_empty[buf], lane);
          ++chunk_ctr;
          fence_proxy_async_global();
          warp_arrive(&sh.a_ready[0], lane);
          if (tid == 128) db_trace(p, 531);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

}  // namespace fcl

namespace fcl {
// Longest-processing-time assignment: tiles are duration-descending; each goes to the currently least-loaded slot
// (ties -> lowest slot). One warp: lane l owns slots l, l+32, ...; the argmin over all slots is ONE redux.sync on the
// key (load << 8 | slot), so an iteration costs tens of cycles (~15 us for 650 tiles). (Assigning whole rounds of
// n_slots tiles at once is faster still but its makespan was 10 % worse on the LJSpeech-shaped batch.)
__global__ void __launch_bounds__(256, 1)
decoder_schedule_kernel(FclDecoderScheduleParams p) {
  extern __shared__ int s_steps[];                    // steps of every tile, loaded in parallel first
  for (int t = threadIdx.x; t < p.n_tiles; t += blockDim.x)
    s_steps[t] = min(max(p.dur[p.order[(size_t)t * p.unit_rows]], 0), FCL_MAX_DURATION);
  __syncthreads();
  if (threadIdx.x >= 32) return;
  constexpr int kPerLane = 8;                         // up to 256 slots
  const int lane = threadIdx.x;
  unsigned key[kPerLane];                             // (load << 8) | slot ; unused slots = max
  int cnt[kPerLane];
#pragma unroll
  for (int i = 0; i < kPerLane; ++i) { key[i] = (lane + 32 * i) < p.n_slots ? (unsigned)(lane + 32 * i) : 0xFFFFFFFFu; cnt[i] = 0; }
  unsigned lmin = key[0];
#pragma unroll
  for (int i = 1; i < kPerLane; ++i) lmin = min(lmin, key[i]);
  for (int t = 0; t < p.n_tiles; ++t) {
    const unsigned best = __reduce_min_sync(0xffffffffu, lmin);
    const int bslot = (int)(best & 0xFFu);
    if ((bslot & 31) == lane) {
      unsigned nm = 0xFFFFFFFFu;
#pragma unroll
      for (int i = 0; i < kPerLane; ++i) {
        if (bslot == lane + 32 * i) {
          key[i] += (unsigned)(s_steps[t] + 1) << 8;
          p.tile_slot[t] = bslot;
          p.tile_rank[t] = cnt[i]++;
        }
        nm = min(nm, key[i]);
      }
      lmin = nm;
    }
  }
}
}  // namespace fcl

extern "C" int fcl_decoder_schedule(const FclDecoderScheduleParams* p, void* stream) {
  using namespace fcl;
  FCL_REQUIRE(p && p->order && p->dur && p->tile_slot && p->tile_rank, "null pointer");
  FCL_REQUIRE(p->unit_rows == 128 || p->unit_rows == 256, "unit_rows must be 128 or 256");
  FCL_REQUIRE(p->n_rows > 0 && p->n_tiles == (p->n_rows + p->unit_rows - 1) / p->unit_rows && p->n_slots >= 1 &&
                  p->n_slots <= 256, "bad sizes");
  FCL_REQUIRE(p->n_tiles <= 12000, "too many tiles for the single-CTA scheduler");
  decoder_schedule_kernel<<<1, 256, (size_t)p->n_tiles * sizeof(int), as_stream(stream)>>>(*p);
  return check_launch("fcl_decoder_schedule");
}

extern "C" int fcl_decoder_bf16_workspace(int32_t prenet_units, int32_t dunits, int64_t* priv_bytes_per_cta,
                                          int64_t* shared_bytes_per_group, int64_t* c_floats_per_cta) {
  if (!priv_bytes_per_cta || !shared_bytes_per_group || !c_floats_per_cta) return FCL_EINVAL;
  *priv_bytes_per_cta = (int64_t)fcl::db_priv_bytes(prenet_units);
  *shared_bytes_per_group = (int64_t)fcl::db_shared_bytes(dunits);
  *c_floats_per_cta = (int64_t)2 * dunits * 128;
  return FCL_OK;
}

extern "C" int fcl_decoder_bf16(const FclDecoderBf16Params* p, void* stream) {
  using namespace fcl;
  FCL_REQUIRE(p && p->order && p->dur && p->frame_off && p->row_utt && p->row_phone && p->hn_img &&
                  p->w_stream && p->bp0 && p->bp1 && p->wpos && p->b0 && p->b1 && p->act_priv && p->act_shared &&
                  p->c_ws && p->before && p->tile_slot && p->tile_rank && p->group_sync,
              "null pointer");
  FCL_REQUIRE(p->group >= 1 && p->group <= 16 && p->n_slots % p->group == 0, "group must divide n_slots (1..16)");
  FCL_REQUIRE((long long)p->n_tiles <= (long long)kDbMaxTilesPerCta * (p->n_slots / p->group), "too many tiles for the per-CTA tile list");
  FCL_REQUIRE(p->eunits % 64 == 0 && p->eunits >= 64, "eunits must be a multiple of 64");
  FCL_REQUIRE(p->n_rows > 0 && p->n_tiles == (p->n_rows + 127) / 128, "n_tiles must be ceil(n_rows / 128)");
  FCL_REQUIRE(p->prenet_units == 256, "prenet_units must be 256 (one 256-column chunk)");
  FCL_REQUIRE(p->dunits % 64 == 0 && p->dunits >= 64, "dunits must be a multiple of 64");
  FCL_REQUIRE(p->odim % 16 == 0 && p->odim <= 128, "odim must be a multiple of 16, <= 128");
  FCL_REQUIRE(p->n_slots >= 1, "n_slots must be >= 1");
  FCL_REQUIRE(p->zoneout >= 0.f && p->zoneout < 1.f && p->dropout_p >= 0.f && p->dropout_p < 1.f, "bad rates");
  const size_t smem = (size_t)kDbStages * kStageBytes;
  if (int rc = ensure_dyn_smem(decoder_bf16_kernel, smem, "fcl_decoder_bf16")) return rc;
  // n_slots CTAs = n_slots / group groups; the spin barriers of group mode need every CTA resident: one CTA per SM,
  // n_slots <= SM count (checked by the caller against fcl_sm_count()).
  cudaError_t e = cudaMemsetAsync(p->group_sync, 0, sizeof(int32_t) * 2 * (size_t)(p->n_slots / p->group), as_stream(stream));
  if (e != cudaSuccess) { set_error("fcl_decoder_bf16: %s", cudaGetErrorString(e)); return FCL_ECUDA; }
  decoder_bf16_kernel<<<p->n_slots, kDbThreads, smem, as_stream(stream)>>>(*p);
  return check_launch("fcl_decoder_bf16");
}
