// K0 -- length regulator: durations -> frame offsets, frame index map, duration-sorted row order.
// Replaces the Python loops of e2e_tts_tacotron2_sa.py:665-671 and decoder_sa.py:619-630 (reference).
// Integer scan/gather work: HBM/L2-latency bound, bit-exact.
#include "common.cuh"

namespace fcl {

constexpr int kScanThreads = 1024;
constexpr int kBins = FCL_MAX_DURATION + 1;   // 1024

__device__ __forceinline__ int warp_incl_scan(int v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int n = __shfl_up_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) >= o) v += n;
  }
  return v;
}

// block-wide exclusive scan of one int per thread (1024 threads); returns exclusive prefix, total in *total
__device__ __forceinline__ int block_excl_scan(int v, int* warp_sums, int* total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int incl = warp_incl_scan(v);
  if (lane == 31) warp_sums[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int s = warp_sums[lane];
    int si = warp_incl_scan(s);
    warp_sums[lane] = si - s;          // exclusive prefix of warp sums
    if (lane == 31) *total = si;
  }
  __syncthreads();
  return incl - v + warp_sums[wid];
}

__global__ void __launch_bounds__(kScanThreads, 1)
len_reg_scan_kernel(FclLenRegParams p) {
  __shared__ int warp_sums[32];
  __shared__ int total_s;
  __shared__ int hist[kBins];
  __shared__ int maxd_s;
  const int tid = threadIdx.x;
  const int P = p.n_rows;
  for (int i = tid; i < kBins; i += kScanThreads) hist[i] = 0;
  if (tid == 0) maxd_s = 0;
  __syncthreads();

  // each thread owns a contiguous chunk of rows
  const int chunk = (P + kScanThreads - 1) / kScanThreads;
  const int lo = min(tid * chunk, P), hi = min(lo + chunk, P);
  int sum = 0, mx = 0;
  for (int k = 0; k < chunk; ++k) {                    // warp-uniform trip count: match_any needs converged lanes
    const int r = lo + k;
    const bool ok = r < hi;
    const int d = ok ? min(max(p.dur[r], 0), FCL_MAX_DURATION) : -1;
    sum += ok ? d : 0;
    mx = max(mx, d);
    // warp-aggregated histogram update: one shared-memory atomic per distinct duration in the warp
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    if (ok && (int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&hist[d], __popc(peers));
  }
  atomicMax(&maxd_s, mx);
  int excl = block_excl_scan(sum, warp_sums, &total_s);
  int run = excl;
  for (int r = lo; r < hi; ++r) {
    p.frame_off[r] = run;
    run += min(max(p.dur[r], 0), FCL_MAX_DURATION);
  }
  if (tid == 0) {
    p.frame_off[P] = total_s;
    p.totals[0] = total_s;
    p.totals[1] = maxd_s;
  }
  __syncthreads();   // frame_off (global, written by this CTA) visible to the CTA

  for (int b = tid; b <= p.n_utts; b += kScanThreads) p.utt_frame_off[b] = p.frame_off[p.utt_off[b]];

  // counting sort by duration, descending: start[k] = #rows with key > k
  // hist has 1024 bins == kScanThreads: one bin per thread, reversed index
  int mybin = kBins - 1 - tid;
  int cnt = hist[mybin];
  __syncthreads();
  int start = block_excl_scan(cnt, warp_sums, &total_s);
  hist[mybin] = start;                 // now a cursor
  __syncthreads();
  for (int k = 0; k < chunk; ++k) {
    const int r = lo + k;
    const bool ok = r < hi;
    const int d = ok ? min(max(p.dur[r], 0), FCL_MAX_DURATION) : -1;
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const int leader = __ffs(peers) - 1, lane = threadIdx.x & 31;
    int base = 0;
    if (ok && leader == lane) base = atomicAdd(&hist[d], __popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (ok) p.order[base + __popc(peers & ((1u << lane) - 1u))] = r;
  }
}

// ---------------------------------------------------------------- multi-CTA form (large batches)
// ws layout (int32): [0, 1024) global histogram | [1024, 2048) per-duration cursors | [2048] max duration |
// [2052, 2052 + G) per-block sums. A block owns kMcRows consecutive rows.
constexpr int kMcRows = 4096;                           // 1024 threads x 4 rows (one int4 load each)

__device__ __forceinline__ int clamp_dur(int d) { return min(max(d, 0), FCL_MAX_DURATION); }

__global__ void __launch_bounds__(kScanThreads, 1)
len_reg_mc_pass1(FclLenRegParams p) {
  __shared__ int warp_sums[32];
  __shared__ int total_s;
  __shared__ int hist[kBins];
  const int tid = threadIdx.x, P = p.n_rows;
  for (int i = tid; i < kBins; i += kScanThreads) hist[i] = 0;
  __syncthreads();
  const int r0 = blockIdx.x * kMcRows + tid * 4;
  int d[4], sum = 0, mx = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    d[j] = r0 + j < P ? clamp_dur(p.dur[r0 + j]) : -1;
    sum += max(d[j], 0);
    mx = max(mx, d[j]);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {                          // warp-aggregated shared-memory histogram
    const unsigned peers = __match_any_sync(0xffffffffu, d[j]);
    if (d[j] >= 0 && (int)(__ffs(peers) - 1) == (tid & 31)) atomicAdd(&hist[d[j]], __popc(peers));
  }
  block_excl_scan(sum, warp_sums, &total_s);
  if (tid == 0) p.ws[2052 + blockIdx.x] = total_s;
  mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 16)); mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
  mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 4)); mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
  mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
  if ((tid & 31) == 0 && mx > 0) atomicMax(&p.ws[2048], mx);
  __syncthreads();
  for (int i = tid; i < kBins; i += kScanThreads)
    if (hist[i]) atomicAdd(&p.ws[i], hist[i]);
}

__global__ void __launch_bounds__(kScanThreads, 1)
len_reg_mc_pass2(FclLenRegParams p, int n_blocks) {
  __shared__ int warp_sums[32];
  __shared__ int total_s;
  __shared__ int hist[kBins];                            // local counts, then this block's base per duration
  __shared__ int local_off[kMcRows];                     // exclusive frame offset of every row of the block (block-relative)
  __shared__ int base_s;
  const int tid = threadIdx.x, P = p.n_rows, lo = blockIdx.x * kMcRows, hi = min(lo + kMcRows, P);
  for (int i = tid; i < kBins; i += kScanThreads) hist[i] = 0;
  // frames before this block = sum of the earlier blocks' sums (a few dozen values)
  int part = 0;
  for (int b = tid; b < (int)blockIdx.x; b += kScanThreads) part += p.ws[2052 + b];
  block_excl_scan(part, warp_sums, &total_s);
  if (tid == 0) base_s = total_s;
  __syncthreads();
  const int base = base_s;
  const int r0 = lo + tid * 4;
  int d[4], sum = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) { d[j] = r0 + j < P ? clamp_dur(p.dur[r0 + j]) : -1; sum += max(d[j], 0); }
  int run = block_excl_scan(sum, warp_sums, &total_s);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (d[j] >= 0) { p.frame_off[r0 + j] = base + run; local_off[tid * 4 + j] = run; run += d[j]; }
  }
  if (blockIdx.x == n_blocks - 1 && tid == 0) {
    p.frame_off[P] = base + total_s;
    p.totals[0] = base + total_s;
    p.totals[1] = p.ws[2048];
  }
  // local histogram + rank of every row among the block's rows of the same duration
  int rank[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const unsigned peers = __match_any_sync(0xffffffffu, d[j]);
    const int leader = __ffs(peers) - 1, lane = tid & 31;
    int b = 0;
    if (d[j] >= 0 && leader == lane) b = atomicAdd(&hist[d[j]], __popc(peers));
    b = __shfl_sync(0xffffffffu, b, leader);
    rank[j] = b + __popc(peers & ((1u << lane) - 1u));
  }
  __syncthreads();
  // per-utterance frame offsets of the utterances that start in this block
  {
    int ulo = 0, uhi = p.n_utts + 1;                     // first b with utt_off[b] >= lo
    while (ulo < uhi) { const int mid = (ulo + uhi) >> 1; if (p.utt_off[mid] < lo) ulo = mid + 1; else uhi = mid; }
    for (int b = ulo + tid; b <= p.n_utts; b += kScanThreads) {
      const int r = p.utt_off[b];
      if (r >= hi && !(r == P && blockIdx.x == n_blocks - 1)) break;
      p.utt_frame_off[b] = r < hi ? base + local_off[r - lo] : base + total_s;
    }
  }
  // counting sort by duration, descending: global start of a duration = #rows with a larger one (from the global
  // histogram), plus the range this block reserves on the per-duration cursor
  const int mybin = kBins - 1 - tid;                     // one (reversed) bin per thread
  const int cnt_global = p.ws[mybin];
  const int start = block_excl_scan(cnt_global, warp_sums, &total_s);
  const int mine = hist[mybin];
  __syncthreads();
  hist[mybin] = start + (mine ? atomicAdd(&p.ws[kBins + mybin], mine) : 0);
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (d[j] >= 0) p.order[hist[d[j]] + rank[j]] = r0 + j;
}

__global__ void __launch_bounds__(256)
frame_map_kernel(FclFrameMapParams p) {
  const int F = p.n_frames;
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < F; f += gridDim.x * blockDim.x) {
    // row = last r with frame_off[r] <= f (upper_bound - 1); zero-duration rows are skipped naturally
    int lo = 0, hi = p.n_rows;            // search in frame_off[0..P], answer in [0, P-1]
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (__ldg(&p.frame_off[mid]) <= f) lo = mid; else hi = mid;
    }
    const int row = lo;
    const int base = __ldg(&p.frame_off[row]);
    const int d = __ldg(&p.frame_off[row + 1]) - base;
    const int step = f - base;
    int ulo = 0, uhi = p.n_utts;
    while (uhi - ulo > 1) {
      int mid = (ulo + uhi) >> 1;
      if (__ldg(&p.utt_frame_off[mid]) <= f) ulo = mid; else uhi = mid;
    }
    p.frame_row[f] = row;
    p.frame_step[f] = step;
    p.frame_seg_lo[f] = __ldg(&p.utt_frame_off[ulo]);
    p.frame_seg_hi[f] = __ldg(&p.utt_frame_off[ulo + 1]);
    if (p.position) p.position[f] = __fdiv_rn((float)step, (float)d);
  }
}

}  // namespace fcl

extern "C" int fcl_len_reg_scan(const FclLenRegParams* p, void* stream) {
  using namespace fcl;
  FCL_REQUIRE(p && p->dur && p->utt_off && p->frame_off && p->utt_frame_off && p->order && p->totals, "null pointer");
  FCL_REQUIRE(p->n_rows > 0 && p->n_utts > 0, "empty batch");
  if (p->ws && p->n_rows >= 2 * kMcRows) {
    const int n_blocks = (p->n_rows + kMcRows - 1) / kMcRows;
    cudaError_t e = cudaMemsetAsync(p->ws, 0, sizeof(int32_t) * 2052, as_stream(stream));
    if (e != cudaSuccess) { set_error("fcl_len_reg_scan: %s", cudaGetErrorString(e)); return FCL_ECUDA; }
    len_reg_mc_pass1<<<n_blocks, kScanThreads, 0, as_stream(stream)>>>(*p);
    len_reg_mc_pass2<<<n_blocks, kScanThreads, 0, as_stream(stream)>>>(*p, n_blocks);
    return check_launch("fcl_len_reg_scan");
  }
  len_reg_scan_kernel<<<1, kScanThreads, 0, as_stream(stream)>>>(*p);
  return check_launch("fcl_len_reg_scan");
}

extern "C" int fcl_len_reg_ws_ints(int32_t n_rows) {
  return 2052 + (n_rows + fcl::kMcRows - 1) / fcl::kMcRows + 4;
}

extern "C" int fcl_len_reg_frame_map(const FclFrameMapParams* p, void* stream) {
  using namespace fcl;
  FCL_REQUIRE(p && p->frame_off && p->utt_frame_off && p->frame_row && p->frame_step && p->frame_seg_lo &&
                  p->frame_seg_hi, "null pointer");
  FCL_REQUIRE(p->n_rows > 0 && p->n_utts > 0 && p->n_frames >= 0, "bad sizes");
  if (p->n_frames == 0) return FCL_OK;
  int sms = fcl_sm_count();
  if (sms < 0) return sms;
  int blocks = min((p->n_frames + 255) / 256, sms * 8);
  frame_map_kernel<<<blocks, 256, 0, as_stream(stream)>>>(*p);
  return check_launch("fcl_len_reg_frame_map");
}
