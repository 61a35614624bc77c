// Row-wise kernels of the variance/duration predictors and the pitch/energy embedding add.
//   fcl_layernorm_f32 : espnet LayerNorm over channels (eps 1e-12) + optional Linear(C,1) head
//                       + optional duration rounding (variance_predictor.py:62,66,90; espnet
//                       DurationPredictor.inference, un-vendored).
//   fcl_embed_add_f32 : Conv1d(1,E,k9) of the predicted pitch / energy + add to h
//                       (e2e_tts_tacotron2_sa.py:435-443,657-658; decoder_sa.py:570-571).
// HBM-bound: one warp per row, float4 accesses.
#include "common.cuh"
#include "umma.cuh"
#include <cuda_bf16.h>

namespace fcl {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int kLnMaxPerLane = 8;   // float4 per lane: chans <= 32*4*8 = 1024

__global__ void __launch_bounds__(256)
layernorm_kernel(FclLayerNormParams p) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int c4 = p.chans >> 2;
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < p.rows; r += gridDim.x * warps_per_block) {
    const float4* x = reinterpret_cast<const float4*>(p.x + (size_t)r * p.chans);
    float4 v[kLnMaxPerLane];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kLnMaxPerLane; ++i) {
      const int c = lane + i * 32;
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < c4) { v[i] = __ldg(x + c); s += (v[i].x + v[i].y) + (v[i].z + v[i].w); }
    }
    const float mean = warp_sum(s) / (float)p.chans;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kLnMaxPerLane; ++i) {
      const int c = lane + i * 32;
      if (c < c4) {
        const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
        q += (a * a + b * b) + (cc * cc + d * d);
      }
    }
    const float var = warp_sum(q) / (float)p.chans;       // biased variance (torch layer_norm)
    const float rstd = 1.0f / sqrtf(var + 1e-12f);
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < kLnMaxPerLane; ++i) {
      const int c = lane + i * 32;
      if (c < c4) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma) + c);
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.beta) + c);
        float4 y;
        y.x = (v[i].x - mean) * rstd * g.x + b.x;
        y.y = (v[i].y - mean) * rstd * g.y + b.y;
        y.z = (v[i].z - mean) * rstd * g.z + b.z;
        y.w = (v[i].w - mean) * rstd * g.w + b.w;
        if (p.y) reinterpret_cast<float4*>(p.y + (size_t)r * p.chans)[c] = y;
        if (p.head_w) {
          const float4 w = __ldg(reinterpret_cast<const float4*>(p.head_w) + c);
          dot += (y.x * w.x + y.y * w.y) + (y.z * w.z + y.w * w.w);
        }
      }
    }
    if (p.head_w) {
      dot = warp_sum(dot) + p.head_b;
      if (lane == 0) {
        if (p.head_out) p.head_out[r] = dot;
        if (p.dur_out) {
          // clamp(round_half_even(exp(x) - 1), 0, cap); rintf rounds half to even like torch.round
          float d = rintf(expf(dot) - 1.0f);
          d = fminf(fmaxf(d, 0.f), (float)FCL_MAX_DURATION);
          p.dur_out[r] = (int)d;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256)
embed_add_kernel(FclEmbedAddParams p) {
  // weights staged in shared memory as [tap][chan] (both embeds), one thread per (row, 4 channels)
  extern __shared__ __align__(16) float sw[];
  float* swp = sw;
  float* swe = sw + (size_t)p.taps * p.chans;
  for (int i = threadIdx.x; i < p.taps * p.chans; i += blockDim.x) {
    const int j = i / p.chans, c = i - j * p.chans;
    swp[i] = p.wp[(size_t)c * p.taps + j];
    swe[i] = p.we[(size_t)c * p.taps + j];
  }
  __syncthreads();
  const int c4 = p.chans >> 2;
  const size_t total = (size_t)p.rows * c4;
  const int half = p.taps >> 1;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / c4), c = (int)(i - (size_t)r * c4) * 4;
    const int lo = p.seg_lo[r], hi = p.seg_hi[r];
    float4 acc = __ldg(reinterpret_cast<const float4*>(p.h + (size_t)r * p.chans + c));
    float4 ap = make_float4(0.f, 0.f, 0.f, 0.f), ae = ap;
    for (int j = 0; j < p.taps; ++j) {
      const int src = r + j - half;
      if (src < lo || src >= hi) continue;
      const float pv = __ldg(p.pitch + src), ev = __ldg(p.energy + src);
      const float4 wpv = *reinterpret_cast<const float4*>(swp + (size_t)j * p.chans + c);
      const float4 wev = *reinterpret_cast<const float4*>(swe + (size_t)j * p.chans + c);
      ap.x = fmaf(wpv.x, pv, ap.x); ap.y = fmaf(wpv.y, pv, ap.y); ap.z = fmaf(wpv.z, pv, ap.z); ap.w = fmaf(wpv.w, pv, ap.w);
      ae.x = fmaf(wev.x, ev, ae.x); ae.y = fmaf(wev.y, ev, ae.y); ae.z = fmaf(wev.z, ev, ae.z); ae.w = fmaf(wev.w, ev, ae.w);
    }
    const float4 bp = __ldg(reinterpret_cast<const float4*>(p.bp + c));
    const float4 be = __ldg(reinterpret_cast<const float4*>(p.be + c));
    // same association as the reference: (h + p_emb) + e_emb, embeds carry their bias
    acc.x = (acc.x + (ap.x + bp.x)) + (ae.x + be.x);
    acc.y = (acc.y + (ap.y + bp.y)) + (ae.y + be.y);
    acc.z = (acc.z + (ap.z + bp.z)) + (ae.z + be.z);
    acc.w = (acc.w + (ap.w + bp.w)) + (ae.w + be.w);
    *reinterpret_cast<float4*>(p.hn + (size_t)r * p.chans + c) = acc;
  }
}

// Same arithmetic, but the result goes straight into the decoder's operand image in duration-sorted tile order:
// hn never exists in fp32. A WARP owns 8 consecutive sorted rows (one 128-byte run of every 8-channel slab of the image):
// lane = slab, so h[row] is one coalesced 1 KB read, the row's 18 pitch / energy taps are fetched once by lanes 0-17 and
// broadcast with shuffles, and each lane finally writes its own 8 x 16 contiguous bytes. No shared-memory staging, no
// block barrier in the loop.
__global__ void __launch_bounds__(256)
embed_add_pack_kernel(FclEmbedAddParams p) {
  extern __shared__ __align__(16) float sw[];
  float* swp = sw;
  float* swe = sw + (size_t)p.taps * p.chans;
  for (int i = threadIdx.x; i < p.taps * p.chans; i += blockDim.x) {
    const int j = i / p.chans, c = i - j * p.chans;
    swp[i] = p.wp[(size_t)c * p.taps + j];
    swe[i] = p.we[(size_t)c * p.taps + j];
  }
  __syncthreads();
  const int c8 = p.chans >> 3, taps = p.taps, half = taps >> 1;
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5, gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_groups = ((p.rows + 127) >> 7) * 16;                 // groups of 8 sorted rows (tiles are padded to 128)
  for (int grp = gw; grp < n_groups; grp += warps) {
    // lane l < 8: metadata of row l of the group
    int my_r = -1, my_lo = 0, my_hi = 0;
    if (lane < 8 && grp * 8 + lane < p.rows) { my_r = p.order[grp * 8 + lane]; my_lo = p.seg_lo[my_r]; my_hi = p.seg_hi[my_r]; }
    const int tile = grp >> 4, ri0 = (grp & 15) * 8;
    for (int kc0 = 0; kc0 < c8; kc0 += 32) {
      const int kc = kc0 + lane, c = kc * 8;
      const bool mine = kc < c8;
      uint4* dst = reinterpret_cast<uint4*>(p.img) + ((size_t)tile * c8 + (mine ? kc : 0)) * 128 + ri0;
      // four rows at a time share every weight read: the 18 KB of embedding weights would otherwise be re-read from
      // shared memory for each row (measured: the kernel was shared-memory-bandwidth bound, 82 % L1TEX throughput)
#pragma unroll 1
      for (int q0 = 0; q0 < 8; q0 += 4) {
        int rq[4];
        float tv[4], ap[4][8], ae[4][8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int r = __shfl_sync(0xffffffffu, my_r, q0 + q), lo = __shfl_sync(0xffffffffu, my_lo, q0 + q),
                    hi = __shfl_sync(0xffffffffu, my_hi, q0 + q);
          rq[q] = r;
          // lanes 0 .. taps-1: pitch taps, lanes 16 .. 16+taps-1: energy taps (zero outside the utterance)
          const int j = lane & 15, src = r + j - half;
          tv[q] = (r >= 0 && j < taps && src >= lo && src < hi) ? __ldg((lane < 16 ? p.pitch : p.energy) + src) : 0.f;
#pragma unroll
          for (int k = 0; k < 8; ++k) { ap[q][k] = 0.f; ae[q][k] = 0.f; }
        }
        for (int j = 0; j < taps; ++j) {
          float4 wp0 = make_float4(0.f, 0.f, 0.f, 0.f), wp1 = wp0, we0 = wp0, we1 = wp0;
          if (mine) {
            wp0 = *reinterpret_cast<const float4*>(swp + (size_t)j * p.chans + c); wp1 = *reinterpret_cast<const float4*>(swp + (size_t)j * p.chans + c + 4);
            we0 = *reinterpret_cast<const float4*>(swe + (size_t)j * p.chans + c); we1 = *reinterpret_cast<const float4*>(swe + (size_t)j * p.chans + c + 4);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float pv = __shfl_sync(0xffffffffu, tv[q], j), ev = __shfl_sync(0xffffffffu, tv[q], 16 + j);
            ap[q][0] = fmaf(wp0.x, pv, ap[q][0]); ap[q][1] = fmaf(wp0.y, pv, ap[q][1]); ap[q][2] = fmaf(wp0.z, pv, ap[q][2]); ap[q][3] = fmaf(wp0.w, pv, ap[q][3]);
            ap[q][4] = fmaf(wp1.x, pv, ap[q][4]); ap[q][5] = fmaf(wp1.y, pv, ap[q][5]); ap[q][6] = fmaf(wp1.z, pv, ap[q][6]); ap[q][7] = fmaf(wp1.w, pv, ap[q][7]);
            ae[q][0] = fmaf(we0.x, ev, ae[q][0]); ae[q][1] = fmaf(we0.y, ev, ae[q][1]); ae[q][2] = fmaf(we0.z, ev, ae[q][2]); ae[q][3] = fmaf(we0.w, ev, ae[q][3]);
            ae[q][4] = fmaf(we1.x, ev, ae[q][4]); ae[q][5] = fmaf(we1.y, ev, ae[q][5]); ae[q][6] = fmaf(we1.z, ev, ae[q][6]); ae[q][7] = fmaf(we1.w, ev, ae[q][7]);
          }
        }
        if (mine) {
          const float4 bp0 = __ldg(reinterpret_cast<const float4*>(p.bp + c)), bp1 = __ldg(reinterpret_cast<const float4*>(p.bp + c) + 1);
          const float4 be0 = __ldg(reinterpret_cast<const float4*>(p.be + c)), be1 = __ldg(reinterpret_cast<const float4*>(p.be + c) + 1);
          const float bpv[8] = {bp0.x, bp0.y, bp0.z, bp0.w, bp1.x, bp1.y, bp1.z, bp1.w};
          const float bev[8] = {be0.x, be0.y, be0.z, be0.w, be1.x, be1.y, be1.z, be1.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 w = make_uint4(0u, 0u, 0u, 0u);
            if (rq[q] >= 0) {
              const float4 h0 = __ldg(reinterpret_cast<const float4*>(p.h + (size_t)rq[q] * p.chans + c));
              const float4 h1 = __ldg(reinterpret_cast<const float4*>(p.h + (size_t)rq[q] * p.chans + c) + 1);
              float acc[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
              for (int k = 0; k < 8; ++k)  // same association as the reference: (h + p_emb) + e_emb, embeds carry their bias
                acc[k] = (acc[k] + (ap[q][k] + bpv[k])) + (ae[q][k] + bev[k]);
              if (p.hn) {
                float4* o = reinterpret_cast<float4*>(p.hn + (size_t)rq[q] * p.chans + c);
                o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
                o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
              }
              w = make_uint4(umma::pack_op(acc[0], acc[1]), umma::pack_op(acc[2], acc[3]), umma::pack_op(acc[4], acc[5]), umma::pack_op(acc[6], acc[7]));
            }
            dst[q0 + q] = w;
          }
        }
      }
    }
  }
}

// gather rows by `order`, round to bf16, write the UMMA operand image [tile][cols/8][128][8]
__global__ void __launch_bounds__(256)
pack_rows_bf16_kernel(FclPackRowsParams p) {
  const int c8 = p.cols >> 3;
  const int tiles = (p.n_rows + 127) >> 7;
  const size_t total = (size_t)tiles * c8 * 128;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i & 127);
    const size_t tk = i >> 7;
    const int kc = (int)(tk % c8), tile = (int)(tk / c8);
    const int sidx = tile * 128 + r;
    uint4 w = make_uint4(0u, 0u, 0u, 0u);
    if (sidx < p.n_rows) {
      const float4* src = reinterpret_cast<const float4*>(p.src + (size_t)p.order[sidx] * p.ld + kc * 8);
      const float4 a = __ldg(src), b = __ldg(src + 1);
      w = make_uint4(umma::pack_op(a.x, a.y), umma::pack_op(a.z, a.w), umma::pack_op(b.x, b.y), umma::pack_op(b.z, b.w));
    }
    reinterpret_cast<uint4*>(p.dst)[i] = w;
  }
}

// prenet.0 of ground-truth frames for the teacher-forced decoder: one thread = (frame, 8 units). The 80 x 256 weight
// is read through L1 (every thread of a warp reads the same k, consecutive units).
__global__ void __launch_bounds__(256)
prenet0_tf_kernel(FclPrenet0TfParams p) {
  const int U = p.prenet_units, O = p.odim, u8 = U >> 3;
  const size_t total = (size_t)p.n_frames * u8;
  const bool use_drop = p.dropout_p > 0.f;
  const uint32_t thr = dropout_threshold16(p.dropout_p);
  const float scale = use_drop ? 1.0f / (1.0f - p.dropout_p) : 1.0f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int f = (int)(i / u8), oct = (int)(i - (size_t)f * u8);
    float acc[8];
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bp0 + 8 * oct)), b1 = __ldg(reinterpret_cast<const float4*>(p.bp0 + 8 * oct) + 1);
    acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w; acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
    const float* yr = p.y + (size_t)f * O;
    for (int k = 0; k < O; ++k) {
      const float yv = __ldg(yr + k);
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.wp0 + (size_t)k * U + 8 * oct));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(p.wp0 + (size_t)k * U + 8 * oct) + 1);
      acc[0] = fmaf(yv, w0.x, acc[0]); acc[1] = fmaf(yv, w0.y, acc[1]); acc[2] = fmaf(yv, w0.z, acc[2]); acc[3] = fmaf(yv, w0.w, acc[3]);
      acc[4] = fmaf(yv, w1.x, acc[4]); acc[5] = fmaf(yv, w1.y, acc[5]); acc[6] = fmaf(yv, w1.z, acc[6]); acc[7] = fmaf(yv, w1.w, acc[7]);
    }
    Philox4 rnd = Philox4{0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
    if (use_drop) {
      const int row = __ldg(p.frame_row + f);
      rnd = dropout_words(p.dropout_seed, (uint32_t)__ldg(p.row_utt + row), (uint32_t)__ldg(p.row_phone + row),
                          (uint32_t)(__ldg(p.frame_step + f) + 1), 0u, (uint32_t)oct);
    }
    const uint32_t wv[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t u16 = (j & 1) ? (wv[j >> 1] >> 16) : (wv[j >> 1] & 0xFFFFu);
      const float v = fmaxf(acc[j], 0.f) * scale;
      x[j] = u16 >= thr ? v : 0.f;
    }
    reinterpret_cast<uint4*>(p.x1)[i] = make_uint4(umma::pack_op(x[0], x[1]), umma::pack_op(x[2], x[3]), umma::pack_op(x[4], x[5]),
                                                   umma::pack_op(x[6], x[7]));
  }
}

}  // namespace fcl

extern "C" int fcl_prenet0_tf(const FclPrenet0TfParams* p, void* stream) {
  using namespace fcl;
  FCL_REQUIRE(p && p->y && p->frame_row && p->frame_step && p->row_utt && p->row_phone && p->wp0 && p->bp0 && p->x1, "null pointer");
  FCL_REQUIRE(p->n_frames > 0 && p->odim > 0 && p->prenet_units % 8 == 0 && p->dropout_p >= 0.f && p->dropout_p < 1.f, "bad sizes");
  int sms = fcl_sm_count();
  if (sms < 0) return sms;
  const size_t total = (size_t)p->n_frames * (p->prenet_units / 8);
  const size_t blocks = (total + 255) / 256;
  prenet0_tf_kernel<<<(unsigned)(blocks < (size_t)sms * 16 ? blocks : (size_t)sms * 16), 256, 0, as_stream(stream)>>>(*p);
  return check_launch("fcl_prenet0_tf");
}

extern "C" int fcl_pack_rows_bf16(const FclPackRowsParams* p, void* stream) {
  using namespace fcl;
  FCL_REQUIRE(p && p->src && p->order && p->dst, "null pointer");
  FCL_REQUIRE(p->n_rows > 0 && p->cols % 8 == 0 && p->ld % 4 == 0, "bad sizes");
  int sms = fcl_sm_count();
  if (sms < 0) return sms;
  const size_t total = (size_t)((p->n_rows + 127) / 128) * (p->cols / 8) * 128;
  int blocks = (int)min((total + 255) / 256, (size_t)sms * 8);
  pack_rows_bf16_kernel<<<blocks, 256, 0, as_stream(stream)>>>(*p);
  return check_launch("fcl_pack_rows_bf16");
}

extern "C" int fcl_layernorm_f32(const FclLayerNormParams* p, void* stream) {
  using namespace fcl;
  FCL_REQUIRE(p && p->x && p->gamma && p->beta, "null pointer");
  FCL_REQUIRE(p->rows > 0 && p->chans > 0 && p->chans % 4 == 0 && p->chans <= 128 * kLnMaxPerLane, "bad sizes");
  int sms = fcl_sm_count();
  if (sms < 0) return sms;
  int blocks = min((p->rows + 7) / 8, sms * 8);
  layernorm_kernel<<<blocks, 256, 0, as_stream(stream)>>>(*p);
  return check_launch("fcl_layernorm_f32");
}

extern "C" int fcl_embed_add_f32(const FclEmbedAddParams* p, void* stream) {
  using namespace fcl;
  FCL_REQUIRE(p && p->h && p->pitch && p->energy && p->seg_lo && p->seg_hi && p->wp && p->bp && p->we && p->be &&
                  (p->hn || p->img), "null pointer");
  FCL_REQUIRE(p->rows > 0 && p->chans % 4 == 0 && (p->taps & 1), "bad sizes");
  FCL_REQUIRE(!p->img || (p->order && p->chans % 8 == 0), "the packed form needs `order` and chans % 8 == 0");
  int sms = fcl_sm_count();
  if (sms < 0) return sms;
  size_t total = (size_t)p->rows * (p->chans / 4);
  int blocks = (int)min((total + 255) / 256, (size_t)sms * 8);
  const size_t smem = (size_t)2 * p->taps * p->chans * sizeof(float);
  FCL_REQUIRE(smem <= 48 * 1024, "embed weights do not fit the static shared-memory window");
  blocks = min(blocks, sms * 4);
  if (p->img) {
    FCL_REQUIRE(p->taps <= 15, "the packed form supports at most 15 taps");
    const int groups = ((p->rows + 127) / 128) * 16;                 // one warp per group of 8 sorted rows
    embed_add_pack_kernel<<<min((groups + 7) / 8, sms * 8), 256, smem, as_stream(stream)>>>(*p);
    return check_launch("fcl_embed_add_f32");
  }
  embed_add_kernel<<<blocks, 256, smem, as_stream(stream)>>>(*p);
  return check_launch("fcl_embed_add_f32");
}
