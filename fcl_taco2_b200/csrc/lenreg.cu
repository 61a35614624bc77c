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
  len_reg_scan_kernel<<<1, kScanThreads, 0, as_stream(stream)>>>(*p);
  return check_launch("fcl_len_reg_scan");
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
