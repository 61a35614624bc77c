// Encoder BiLSTM recurrence on the tensor cores (tcgen05).
// torch.nn.LSTM(E, E/2, 1, bidirectional) of the reference (nets/modules/encoder_sa.py:96-100,143-146), every
// utterance of the batch independently, packed-sequence semantics (the backward direction starts at the last
// valid phoneme of each utterance).
//
// CTA = (tile of 128 utterances, direction). The batch planner orders utterances longest-first, so a tile
// runs for the length of its first utterance and rows drop out as they finish. Per time step
//   gates[128 x 4H] = gx[t] (input projection, precomputed by fcl_conv_gemm_bf16, bf16) + h[128 x H] W_hh^T
// is one M=128 tcgen05 GEMM in 256-column chunks (gate-interleaved columns, two TMEM accumulator buffers so the
// epilogue of chunk j overlaps the MMAs of chunk j+1). h lives in shared memory as a bf16 UMMA operand image
// (double-buffered, written by the epilogue warps: generic->async proxy fence + mbarrier), the cell state in a
// global fp32 scratch, W_hh streams from L2 through a bulk-copy ring (for H = 128 the ring holds all of it).
// The step chain is latency/MUFU-bound (5 transcendental per cell), not tensor-bound.
//
// Warp roles (640 threads): warp 0 = W_hh bulk-copy producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
// warps 4-19 = epilogue (TMEM lane quarter = warp % 4 -> utterance row; column quarter = (warp-4)/4 -> 16 units).
#include "common.cuh"
#include "umma.cuh"
#include <cuda_fp16.h>

namespace fcl {
using namespace umma;

constexpr int kBlThreads = 640;
constexpr int kBlEpiThreads = 512;
constexpr uint32_t kBlBBytes = 256u * 64u * 2u;          // one W_hh stage: 256 gate columns x 64 k

// 16 consecutive gate-interleaved columns (4 hidden units x i,f,g,o) of the input projection of one phoneme:
// either bf16 rows (P, 8H) or the fp32 column-blocked image [8H/16][gx_rows][16] in the padded row space.
__device__ __forceinline__ void load_gx16(const FclBiLstmBf16Params& p, long row, long prow, int col, float (&g)[16]) {
  if (p.gx_blk && p.gx_blk_half) {
    const uint4* s = reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(p.gx_blk) + ((size_t)(col >> 4) * (size_t)p.gx_rows + (size_t)prow) * 16);
    const uint4 a = __ldg(s), b = __ldg(s + 1);
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) { const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[k])); g[2 * k] = f.x; g[2 * k + 1] = f.y; }
  } else if (p.gx_blk) {
    const float4* s = reinterpret_cast<const float4*>(p.gx_blk + ((size_t)(col >> 4) * (size_t)p.gx_rows + (size_t)prow) * 16);
#pragma unroll
    for (int k = 0; k < 4; ++k) { const float4 v = __ldg(s + k); g[4 * k] = v.x; g[4 * k + 1] = v.y; g[4 * k + 2] = v.z; g[4 * k + 3] = v.w; }
  } else {
    const uint4* s = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.gx) + (size_t)row * (size_t)(8 * p.hidden) + col);
    const uint4 a = __ldg(s), b = __ldg(s + 1);
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) { g[2 * k] = op_lo(w[k]); g[2 * k + 1] = op_hi(w[k]); }
  }
}
__device__ __forceinline__ void prefetch_gx16(const FclBiLstmBf16Params& p, long row, long prow, int col) {
  const void* a = p.gx_blk ? (p.gx_blk_half ? static_cast<const void*>(reinterpret_cast<const __half*>(p.gx_blk) + ((size_t)(col >> 4) * (size_t)p.gx_rows + (size_t)prow) * 16)
                                            : static_cast<const void*>(p.gx_blk + ((size_t)(col >> 4) * (size_t)p.gx_rows + (size_t)prow) * 16))
                           : static_cast<const void*>(reinterpret_cast<const uint16_t*>(p.gx) + (size_t)row * (size_t)(8 * p.hidden) + col);
  asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
}

// one mbarrier arrival per epilogue WARP (the lanes' tcgen05.ld / fenced shared-memory writes are ordered before lane 0's
// arrive by the warp barrier): 16 arrivals per hand-over instead of 512 on the per-step critical path
__device__ __forceinline__ void warp_arrive1(uint64_t* bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

// Counter-mode profile (profiling build -DFCL_DEC_PROF only, tools/bilstm_prof.py): CTA (0, 0) accumulates %clock deltas
// and dumps them into the first words of c_ws (unused by the replicated modes): [0] steps, [1] issuer: waiting for
// h_ready, [2] issuer: issuing a step, [3 + 2 * half] epilogue half: waiting for its accumulator, [4 + 2 * half] its body
#ifdef FCL_DEC_PROF
#define BL_PROF(idx) do { if (prof) { const uint32_t n_ = clk32b(); sh.prof[idx] += n_ - pt; pt = n_; } } while (0)
__device__ __forceinline__ uint32_t clk32b() { uint32_t c; asm volatile("mov.u32 %0, %%clock;" : "=r"(c)); return c; }
#else
#define BL_PROF(idx) do { } while (0)
#endif

struct BlShared {
  uint64_t full[4], empty[4];
  uint64_t tmem_full[2], tmem_empty[2];
  uint64_t h_ready;
  uint32_t tmem_base;
#ifdef FCL_DEC_PROF
  uint32_t prof[8];
#endif
};

__global__ void __launch_bounds__(kBlThreads, 1)
bilstm_bf16_kernel(FclBiLstmBf16Params p, int stages) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ BlShared sh;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = p.hidden, dir = blockIdx.y, tile = blockIdx.x;
  const int kH = H / 64, nch = 4 * H / 256;
  const uint32_t himg_bytes = (uint32_t)H * 256u;         // [H/8][128][8] bf16
  uint8_t* himg = smem + (size_t)stages * kBlBBytes;       // two h images after the ring

  const int R = p.tile_utts;                              // utterances per tile (rows >= R stay idle)
  const int u_first = tile * R;
  __shared__ int s_steps;
  if (tid == 0) s_steps = 0;
  __syncthreads();
  if (tid < R && u_first + tid < p.n_utts) atomicMax(&s_steps, p.utt_off[u_first + tid + 1] - p.utt_off[u_first + tid]);
  __syncthreads();
  const int steps = s_steps;                             // longest utterance of the tile

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&sh.full[s], 1); mbar_init(&sh.empty[s], 1); }
    // split mode (below): a chunk's accumulator is drained by one HALF of the epilogue warps
    const bool split = p.tile_utts == 32 && (nch & 1) == 0;
    for (int b = 0; b < 2; ++b) { mbar_init(&sh.tmem_full[b], 1); mbar_init(&sh.tmem_empty[b], (kBlEpiThreads / 32) / (split ? 2 : 1)); }
    mbar_init(&sh.h_ready, kBlEpiThreads / 32);
    fence_barrier_init();
  }
#ifdef FCL_DEC_PROF
  if (tid < 8) sh.prof[tid] = 0u;
#endif
  if (warp == 2) tmem_alloc(&sh.tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh.tmem_base;

  if (warp == 0) {
    // ================================================================ W_hh producer (independent of h)
    if (elect_one()) {
      uint32_t stage = 0, sphase = 0;
      const uint8_t* wdir = reinterpret_cast<const uint8_t*>(p.whh_packed) + (size_t)dir * nch * kH * kBlBBytes;
      // W_hh RESIDENT (H = 128: the ring holds all nch * kH stages): copied once, never refilled -- the issuer then has no
      // full / empty traffic on its per-step critical path
      const int n_loads = (nch * kH <= stages) ? min(steps, 1) : steps;
      for (int t = 0; t < n_loads; ++t) {
        const uint8_t* wptr = wdir;
        for (int i = 0; i < nch * kH; ++i) {
          mbar_wait(&sh.empty[stage], sphase ^ 1u);
          mbar_arrive_expect_tx(&sh.full[stage], kBlBBytes);
          bulk_g2s(smem + (size_t)stage * kBlBBytes, wptr, kBlBBytes, &sh.full[stage]);
          wptr += kBlBBytes;
          if (++stage == (uint32_t)stages) { stage = 0; sphase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    if (elect_one()) {
      uint32_t stage = 0, sphase = 0, chunk_ctr = 0;
      const uint32_t idesc = idesc_op_f32(128u, 256u);
      const bool resident = nch * kH <= stages;
      // descriptors as in the decoder kernels: high word constant (SBO 128 B, version 1), low word = (address >> 4) |
      // (LBO >> 4) << 16 built with one add per MMA -- this thread's instruction chain is on the step's critical path
      constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);
      constexpr uint32_t kALo = (2048u >> 4) << 16, kBLo = (4096u >> 4) << 16;
      const uint32_t ring_lo = smem_u32(smem) >> 4;
#ifdef FCL_DEC_PROF
      const bool prof = blockIdx.x == 0 && blockIdx.y == 0;
      uint32_t pt = clk32b();
#endif
      if (resident && nch == 2 && kH == 2) {
        // H = 128: 16 MMAs per step against resident weights, fully unrolled with immediate descriptor offsets. The
        // profile (tools/bilstm_prof.py) showed this thread taking 144 cycles per MMA in the generic loop below: more
        // than the 128 cycles the pipe needs, and all of it on the step's critical chain.
        const uint32_t a0 = (smem_u32(himg) >> 4) + kALo, a_par = himg_bytes >> 4, b0 = ring_lo + kBLo;
        for (int t = 0; t < steps; ++t) {
          if (t == 0) {
            for (int s2 = 0; s2 < 4; ++s2) mbar_wait(&sh.full[s2], 0u);
          }
          mbar_wait(&sh.h_ready, (uint32_t)t & 1u);
          tc_fence_after();
          BL_PROF(1);
          const uint32_t a_lo = a0 + ((uint32_t)t & 1u) * a_par;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            mbar_wait(&sh.tmem_empty[c], ((uint32_t)t & 1u) ^ 1u);   // buffer c is used once per step
            tc_fence_after();
#pragma unroll
            for (int i = 0; i < 8; ++i) {                    // i = ks * 4 + k
              const uint64_t ad = ((uint64_t)kDescHi << 32) | (a_lo + (uint32_t)i * (4096u >> 4));
              const uint64_t bd = ((uint64_t)kDescHi << 32) |
                                  (b0 + (uint32_t)(c * 2 + (i >> 2)) * (kBlBBytes >> 4) + (uint32_t)(i & 3) * (8192u >> 4));
              mma_bf16_ss(tmem + (uint32_t)c * 256u, ad, bd, idesc, i > 0 ? 1u : 0u);
            }
            mma_commit(&sh.tmem_full[c]);
          }
          BL_PROF(2);
        }
      } else
      for (int t = 0; t < steps; ++t) {
        mbar_wait(&sh.h_ready, (uint32_t)t & 1u);          // h(t-1) image complete (t = 0: zeros)
        tc_fence_after();
        BL_PROF(1);
        uint32_t a_lo = (smem_u32(himg + (size_t)(t & 1) * himg_bytes) >> 4) + kALo;
        for (int c = 0; c < nch; ++c) {
          const uint32_t buf = chunk_ctr & 1u, use = chunk_ctr >> 1;
          mbar_wait(&sh.tmem_empty[buf], (use & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem + buf * 256u;
          for (int ks = 0; ks < kH; ++ks) {
            if (!resident || t == 0) {
              mbar_wait(&sh.full[stage], sphase);
              tc_fence_after();
            }
            const uint32_t b_lo = ring_lo + (uint32_t)stage * (kBlBBytes >> 4) + kBLo;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t ad = ((uint64_t)kDescHi << 32) | (a_lo + (uint32_t)k * (4096u >> 4));
              const uint64_t bd = ((uint64_t)kDescHi << 32) | (b_lo + (uint32_t)k * (8192u >> 4));
              mma_bf16_ss(d_tmem, ad, bd, idesc, (ks > 0 || k > 0) ? 1u : 0u);
            }
            a_lo += 4u * (4096u >> 4);
            if (!resident) mma_commit(&sh.empty[stage]);
            if (++stage == (uint32_t)(resident ? nch * kH : stages)) { stage = 0; sphase ^= 1u; }
          }
          a_lo -= (uint32_t)kH * 4u * (4096u >> 4);
          mma_commit(&sh.tmem_full[buf]);
          ++chunk_ctr;
        }
        BL_PROF(2);
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ================================================================ epilogue
    const int q = warp & 3, cs = (warp - 4) >> 2;
    const int r = q * 32 + lane;                           // utterance row of the tile == TMEM lane
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
    const int u = u_first + r;
    int off = 0, len = 0;
    long poff = 0;
    if (r < R && u < p.n_utts) {
      off = p.utt_off[u]; len = p.utt_off[u + 1] - off;
      if (p.prow_off) poff = p.prow_off[u];
    }
    float* cst = p.c_ws + ((size_t)(tile * 2 + dir) * H) * 128;        // [H][128]
    uint32_t chunk_ctr = 0;

    // zero the first h image (this thread's quarter of the k-chunks)
    for (int kc = cs; kc < H / 8; kc += 4) *reinterpret_cast<uint4*>(himg + ((size_t)kc * 128 + r) * 16) = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
    warp_arrive1(&sh.h_ready, lane);
    // The zeroes above and the h rows written into the same image from step 1 on come from different threads. They are
    // ordered through the issuer (h_ready -> MMAs -> tcgen05.commit -> tmem_full), a path compute-sanitizer's racecheck
    // does not model; one named barrier among the epilogue warps makes the order explicit (once per kernel).
    asm volatile("bar.sync 1, 512;" ::: "memory");

    if (R == 32 && (nch & 1) == 0) {
      // ---- replicated + split mode (even chunk counts: H = 128, 256). As below the 32 utterances of the tile occupy all
      // four 32-row quarters of the M = 128 operand, but the chunks of a step are dealt to the two HALVES of the epilogue
      // warps (chunk c -> half c & 1 == its accumulator buffer): both halves update cells at the same time, each thread
      // 8 hidden units of its chunk, instead of all warps doing 4 units of every chunk one chunk after the other. The
      // step is a latency chain (MMA -> wake -> update -> fence -> arrive -> wake) and the chunk epilogues were its longer
      // part (S batch 1024: 0.42 -> see profiles/r02_bilstm.md).
      const int uu = u_first + lane;
      int off2 = 0, len2 = 0;
      long poff2 = 0;
      if (uu < p.n_utts) {
        off2 = p.utt_off[uu]; len2 = p.utt_off[uu + 1] - off2;
        if (p.prow_off) poff2 = p.prow_off[uu];
      }
      const int half = cs >> 1;                              // which chunks (c & 1 == half) and which accumulator buffer
      const int usub = ((cs & 1) * 4 + q) * 8;               // first of this thread's 8 units within a 64-unit chunk
      float creg[2][8];                                      // chunks half, half + 2
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int j = 0; j < 8; ++j) creg[c][j] = 0.f;
      uint32_t my_use = 0;                                   // uses of accumulator buffer `half` so far
#ifdef FCL_DEC_PROF
      const bool prof = blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && (warp == 4 || warp == 12);
      uint32_t pt = clk32b();
#endif
      for (int t = 0; t < steps; ++t) {
        const bool active = t < len2;
        const int tt = dir == 0 ? t : len2 - 1 - t;
        const int grow = off2 + tt;
        const long gprow = poff2 + tt;
        uint8_t* hnew = himg + (size_t)((t + 1) & 1) * himg_bytes;
#pragma unroll
        for (int ci = 0; ci < 2; ++ci) {
          const int c = half + 2 * ci;
          if (c < nch) {
            const int ub = c * 64 + usub;
            const int col = dir * 4 * H + 4 * ub;            // gate-interleaved column of unit ub, gate i
            float gq[2][16];
#pragma unroll
            for (int j = 0; j < 16; ++j) gq[0][j] = gq[1][j] = 0.f;
            if (active) {                                    // requested before waiting for the accumulator
              load_gx16(p, grow, gprow, col, gq[0]);
              load_gx16(p, grow, gprow, col + 16, gq[1]);
              if (t + 1 < len2) {                            // pull the next step's gx into L2
                prefetch_gx16(p, grow + (dir == 0 ? 1 : -1), gprow + (dir == 0 ? 1 : -1), col);
                if (p.gx_blk) prefetch_gx16(p, grow + (dir == 0 ? 1 : -1), gprow + (dir == 0 ? 1 : -1), col + 16);
              }
            }
            mbar_wait(&sh.tmem_full[half], my_use & 1u);
            tc_fence_after();
            BL_PROF(3 + 2 * half);
            float hf[8];
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              float v[16];
              tmem_ld16(lane_addr + (uint32_t)half * 256u + (uint32_t)(usub * 4 + g * 16), v);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float ig = sigmoid_fast(v[4 * j] + gq[g][4 * j]), fg = sigmoid_fast(v[4 * j + 1] + gq[g][4 * j + 1]);
                const float cg = tanh_fast(v[4 * j + 2] + gq[g][4 * j + 2]), og = sigmoid_fast(v[4 * j + 3] + gq[g][4 * j + 3]);
                const float cn = fmaf(fg, creg[ci][g * 4 + j], ig * cg);
                hf[g * 4 + j] = active ? og * tanh_fast(cn) : 0.f;
                if (active) creg[ci][g * 4 + j] = cn;
              }
            }
            tc_fence_before();
            warp_arrive1(&sh.tmem_empty[half], lane);
            ++my_use;
            const uint4 hw = make_uint4(pack_op(hf[0], hf[1]), pack_op(hf[2], hf[3]), pack_op(hf[4], hf[5]), pack_op(hf[6], hf[7]));
#pragma unroll
            for (int k = 0; k < 4; ++k)                      // all four replicas of this utterance's row
              *reinterpret_cast<uint4*>(hnew + ((size_t)(ub >> 3) * 128 + k * 32 + lane) * 16) = hw;
            if (active) {
              float4* o = reinterpret_cast<float4*>(p.out + (size_t)grow * 2 * H + (size_t)dir * H + ub);
              o[0] = make_float4(hf[0], hf[1], hf[2], hf[3]);
              o[1] = make_float4(hf[4], hf[5], hf[6], hf[7]);
            }
          }
        }
        fence_proxy_async_smem();
        warp_arrive1(&sh.h_ready, lane);
        BL_PROF(4 + 2 * half);
      }
    } else if (R == 32) {
      // ---- replicated mode: the 32 utterances of the tile occupy all four 32-row quarters of the M = 128 operand
      // (the same h rows four times), so all 16 epilogue warps share the cell updates: thread (quarter q, column
      // set cs, lane) owns utterance `lane` and the 4 hidden units (cs*4 + q)*4.. of every chunk; its cell state
      // stays in registers. The serial chain per step is 4x shorter than with one quarter doing all the work.
      const int uu = u_first + lane;
      int off2 = 0, len2 = 0;
      long poff2 = 0;
      if (uu < p.n_utts) {
        off2 = p.utt_off[uu]; len2 = p.utt_off[uu + 1] - off2;
        if (p.prow_off) poff2 = p.prow_off[uu];
      }
      const int usub = (cs * 4 + q) * 4;                     // first unit within a 64-unit chunk
      float creg[4][4];
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int j = 0; j < 4; ++j) creg[c][j] = 0.f;
      for (int t = 0; t < steps; ++t) {
        const bool active = t < len2;
        const int tt = dir == 0 ? t : len2 - 1 - t;
        const int grow = off2 + tt;
        const long gprow = poff2 + tt;
        uint8_t* hnew = himg + (size_t)((t + 1) & 1) * himg_bytes;
        // input projection of this step, one chunk AHEAD of the cell update: the request for chunk c + 1 is in flight
        // while chunk c waits for its accumulator, so no L2 latency sits between "accumulator ready" and the update
        // (the epilogue is the longer pole of a step: two chunk epilogues against 2 x 1024 tensor cycles)
        float gq[2][16];
#pragma unroll
        for (int j = 0; j < 16; ++j) gq[0][j] = gq[1][j] = 0.f;
        if (active) load_gx16(p, grow, gprow, dir * 4 * H + 4 * usub, gq[0]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c < nch) {
            const int ub = c * 64 + usub;
            const int col = dir * 4 * H + 4 * ub;              // gate-interleaved column of unit ub, gate i
            if (active) {
              if (c + 1 < nch) load_gx16(p, grow, gprow, col + 4 * 64, gq[(c + 1) & 1]);
              if (t + 1 < len2 && (p.gx_blk || (usub & 15) == 0))   // pull the next step's gx into L2
                prefetch_gx16(p, grow + (dir == 0 ? 1 : -1), gprow + (dir == 0 ? 1 : -1), col);
            }
            const uint32_t buf = chunk_ctr & 1u, use = chunk_ctr >> 1;
            mbar_wait(&sh.tmem_full[buf], use & 1u);
            tc_fence_after();
            float v[16], hf[4];
            tmem_ld16(lane_addr + buf * 256u + (uint32_t)(usub * 4), v);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float ig = sigmoid_fast(v[4 * j] + gq[c & 1][4 * j]), fg = sigmoid_fast(v[4 * j + 1] + gq[c & 1][4 * j + 1]);
              const float cg = tanh_fast(v[4 * j + 2] + gq[c & 1][4 * j + 2]), og = sigmoid_fast(v[4 * j + 3] + gq[c & 1][4 * j + 3]);
              const float cn = fmaf(fg, creg[c][j], ig * cg);
              hf[j] = active ? og * tanh_fast(cn) : 0.f;
              if (active) creg[c][j] = cn;
            }
            tc_fence_before();
            warp_arrive1(&sh.tmem_empty[buf], lane);
            ++chunk_ctr;
            const uint2 hw = make_uint2(pack_op(hf[0], hf[1]), pack_op(hf[2], hf[3]));
#pragma unroll
            for (int k = 0; k < 4; ++k)                      // all four replicas of this utterance's row
              *reinterpret_cast<uint2*>(hnew + ((size_t)(ub >> 3) * 128 + k * 32 + lane) * 16 + (ub & 4) * 2) = hw;
            if (active)
              *reinterpret_cast<float4*>(p.out + (size_t)grow * 2 * H + (size_t)dir * H + ub) = make_float4(hf[0], hf[1], hf[2], hf[3]);
          }
        }
        fence_proxy_async_smem();
        warp_arrive1(&sh.h_ready, lane);
      }
    } else
    for (int t = 0; t < steps; ++t) {
      const bool active = t < len;
      const int tt = dir == 0 ? t : len - 1 - t;
      const int grow = off + tt;
      const long gprow = poff + tt;
      uint8_t* hnew = himg + (size_t)((t + 1) & 1) * himg_bytes;
#pragma unroll 1
      for (int c = 0; c < nch; ++c) {
        const int u0 = c * 64 + cs * 16;                   // first of this thread's 16 hidden units
        const int col0 = dir * 4 * H + 4 * u0;
        // request the global operands before waiting for the accumulator
        float gxv[2][16];                                 // two-deep: group g+1 is requested while group g is computed
        float cold[16];
        if (active) {
          load_gx16(p, grow, gprow, col0, gxv[0]);
#pragma unroll
          for (int j = 0; j < 16; ++j) cold[j] = t == 0 ? 0.f : __ldcg(cst + (size_t)(u0 + j) * 128 + r);
          if (t + 1 < len) {                               // pull the next step's gx lines into L2
            prefetch_gx16(p, grow + (dir == 0 ? 1 : -1), gprow + (dir == 0 ? 1 : -1), col0);
            if (p.gx_blk) {                                 // the four 16-column blocks of this thread are separate lines
#pragma unroll
              for (int j = 1; j < 4; ++j) prefetch_gx16(p, grow, gprow + (dir == 0 ? 1 : -1), col0 + 16 * j);
            }
          }
        }
        const uint32_t buf = chunk_ctr & 1u, use = chunk_ctr >> 1;
        mbar_wait(&sh.tmem_full[buf], use & 1u);
        tc_fence_after();
        uint32_t hout[8];
        float hf[16];
#pragma unroll
        for (int g = 0; g < 4; ++g) {                      // 4 units (16 accumulator columns) at a time
          float v[16];
          if (active && g + 1 < 4) load_gx16(p, grow, gprow, col0 + 16 * (g + 1), gxv[(g + 1) & 1]);
          tmem_ld16(lane_addr + buf * 256u + (uint32_t)(cs * 64 + g * 16), v);
          if (active) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int ul = g * 4 + j;
              const float gi = gxv[g & 1][4 * j], gf = gxv[g & 1][4 * j + 1], gg = gxv[g & 1][4 * j + 2], go = gxv[g & 1][4 * j + 3];
              const float ig = sigmoid_fast(v[4 * j] + gi), fg = sigmoid_fast(v[4 * j + 1] + gf);
              const float cg = tanh_fast(v[4 * j + 2] + gg), og = sigmoid_fast(v[4 * j + 3] + go);
              const float cn = fmaf(fg, cold[ul], ig * cg);
              hf[ul] = og * tanh_fast(cn);
              cst[(size_t)(u0 + ul) * 128 + r] = cn;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) hf[g * 4 + j] = 0.f;
          }
          hout[2 * g] = pack_op(hf[g * 4], hf[g * 4 + 1]);
          hout[2 * g + 1] = pack_op(hf[g * 4 + 2], hf[g * 4 + 3]);
        }
        tc_fence_before();
        warp_arrive1(&sh.tmem_empty[buf], lane);
        ++chunk_ctr;
        *reinterpret_cast<uint4*>(hnew + ((size_t)(u0 >> 3) * 128 + r) * 16) = make_uint4(hout[0], hout[1], hout[2], hout[3]);
        *reinterpret_cast<uint4*>(hnew + ((size_t)((u0 >> 3) + 1) * 128 + r) * 16) = make_uint4(hout[4], hout[5], hout[6], hout[7]);
        if (active) {
          float4* o = reinterpret_cast<float4*>(p.out + (size_t)grow * 2 * H + (size_t)dir * H + u0);
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] = make_float4(hf[4 * j], hf[4 * j + 1], hf[4 * j + 2], hf[4 * j + 3]);
        }
      }
      fence_proxy_async_smem();
      warp_arrive1(&sh.h_ready, lane);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
#ifdef FCL_DEC_PROF
  if (blockIdx.x == 0 && blockIdx.y == 0 && tid < 8) reinterpret_cast<uint32_t*>(p.c_ws)[tid] = tid == 0 ? (uint32_t)steps : sh.prof[tid];
#endif
}

}  // namespace fcl

extern "C" int fcl_bilstm_bf16(const FclBiLstmBf16Params* p, void* stream) {
  using namespace fcl;
  FCL_REQUIRE(p && p->utt_off && (p->gx || p->gx_blk) && p->whh_packed && p->c_ws && p->out, "null pointer");
  FCL_REQUIRE(!p->gx_blk || (p->prow_off && p->gx_rows > 0), "gx_blk needs prow_off and gx_rows");
  FCL_REQUIRE(p->n_utts > 0, "empty batch");
  FCL_REQUIRE(p->hidden % 64 == 0 && p->hidden >= 64 && p->hidden <= 256, "hidden must be 64, 128, 192 or 256");
  const int stages = p->hidden <= 128 ? 4 : 3;
  const size_t smem = (size_t)stages * kBlBBytes + 2 * (size_t)p->hidden * 256;
  FCL_REQUIRE(smem <= 226 * 1024, "shared memory budget exceeded");
  if (int rc = ensure_dyn_smem(bilstm_bf16_kernel, smem, "fcl_bilstm_bf16")) return rc;
  FCL_REQUIRE(p->tile_utts == 32 || p->tile_utts == 64 || p->tile_utts == 128, "tile_utts must be 32, 64 or 128");
  dim3 grid((p->n_utts + p->tile_utts - 1) / p->tile_utts, 2);
  bilstm_bf16_kernel<<<grid, kBlThreads, smem, as_stream(stream)>>>(*p, stages);
  return check_launch("fcl_bilstm_bf16");
}
