// Encoder BiLSTM recurrence, fp32 CUDA-core path.
// torch.nn.LSTM(E, E/2, 1, bidirectional) (reference: nets/modules/encoder_sa.py:96-100,143-146),
// one CTA per (group of utterances, direction); thread = hidden unit, gates i,f,g,o of that unit are
// adjacent in the packed weights so the cell update stays in registers. W_hh streams from L2.
#include "common.cuh"

namespace fcl {

template <int UB>
__global__ void __launch_bounds__(256)
bilstm_f32_kernel(FclBiLstmParams p) {
  extern __shared__ __align__(16) float hs[];          // [hidden][UB]
  __shared__ int s_off[UB], s_len[UB];
  const int H = p.hidden, u = threadIdx.x, dir = blockIdx.y;
  const int g0 = blockIdx.x * UB;
  if (u < UB) {
    const int b = g0 + u;
    if (b < p.n_utts) { s_off[u] = p.utt_off[b]; s_len[u] = p.utt_off[b + 1] - p.utt_off[b]; }
    else { s_off[u] = 0; s_len[u] = 0; }
  }
  for (int i = u; i < H * UB; i += blockDim.x) hs[i] = 0.f;
  __syncthreads();
  int maxlen = 0;
#pragma unroll
  for (int j = 0; j < UB; ++j) maxlen = max(maxlen, s_len[j]);

  float c[UB];
#pragma unroll
  for (int j = 0; j < UB; ++j) c[j] = 0.f;
  const float* __restrict__ w = p.whh + (size_t)dir * H * 4 * H + 4 * u;
  const size_t gx_ld = (size_t)8 * H;

  for (int step = 0; step < maxlen; ++step) {
    float acc[UB][4];
    int row[UB];
#pragma unroll
    for (int j = 0; j < UB; ++j) {
      const bool act = step < s_len[j];
      row[j] = act ? s_off[j] + (dir == 0 ? step : s_len[j] - 1 - step) : -1;
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      if (act) g = __ldg(reinterpret_cast<const float4*>(p.gx + (size_t)row[j] * gx_ld + (size_t)dir * 4 * H + 4 * u));
      acc[j][0] = g.x; acc[j][1] = g.y; acc[j][2] = g.z; acc[j][3] = g.w;
    }
#pragma unroll 8
    for (int k = 0; k < H; ++k) {
      const float4 w4 = __ldg(reinterpret_cast<const float4*>(w + (size_t)k * 4 * H));
      float hv[UB];
      if (UB == 8) {
        const float4 h0 = *reinterpret_cast<const float4*>(&hs[k * UB]);
        const float4 h1 = *reinterpret_cast<const float4*>(&hs[k * UB + 4]);
        hv[0] = h0.x; hv[1] = h0.y; hv[2] = h0.z; hv[3] = h0.w;
        hv[4 % UB] = h1.x; hv[5 % UB] = h1.y; hv[6 % UB] = h1.z; hv[7 % UB] = h1.w;
      } else {
#pragma unroll
        for (int j = 0; j < UB; ++j) hv[j] = hs[k * UB + j];
      }
#pragma unroll
      for (int j = 0; j < UB; ++j) {
        acc[j][0] = fmaf(w4.x, hv[j], acc[j][0]);
        acc[j][1] = fmaf(w4.y, hv[j], acc[j][1]);
        acc[j][2] = fmaf(w4.z, hv[j], acc[j][2]);
        acc[j][3] = fmaf(w4.w, hv[j], acc[j][3]);
      }
    }
    float hn[UB];
#pragma unroll
    for (int j = 0; j < UB; ++j) {
      const float ig = sigmoid_acc(acc[j][0]), fg = sigmoid_acc(acc[j][1]);
      const float gg = tanhf(acc[j][2]), og = sigmoid_acc(acc[j][3]);
      const float cn = fg * c[j] + ig * gg;
      hn[j] = og * tanhf(cn);
      if (row[j] >= 0) c[j] = cn;
    }
    __syncthreads();                       // everyone finished reading hs
#pragma unroll
    for (int j = 0; j < UB; ++j) {
      if (row[j] >= 0) {
        hs[u * UB + j] = hn[j];
        p.out[(size_t)row[j] * 2 * H + (size_t)dir * H + u] = hn[j];
      }
    }
    __syncthreads();
  }
}

}  // namespace fcl

extern "C" int fcl_bilstm_f32(const FclBiLstmParams* p, void* stream) {
  using namespace fcl;
  FCL_REQUIRE(p && p->utt_off && p->gx && p->whh && p->out, "null pointer");
  FCL_REQUIRE(p->n_utts > 0 && p->hidden >= 32 && p->hidden <= 256 && p->hidden % 32 == 0, "hidden must be 32..256");
  FCL_REQUIRE(p->group == 1 || p->group == 8, "group must be 1 or 8");
  cudaStream_t s = as_stream(stream);
  if (p->group == 8) {
    dim3 grid((p->n_utts + 7) / 8, 2);
    bilstm_f32_kernel<8><<<grid, p->hidden, (size_t)p->hidden * 8 * sizeof(float), s>>>(*p);
  } else {
    dim3 grid(p->n_utts, 2);
    bilstm_f32_kernel<1><<<grid, p->hidden, (size_t)p->hidden * sizeof(float), s>>>(*p);
  }
  return check_launch("fcl_bilstm_f32");
}
