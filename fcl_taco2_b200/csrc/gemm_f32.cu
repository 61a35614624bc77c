// Ragged conv1d / linear as an fp32 GEMM on CUDA cores (the fp32 parity path).
// out[r,n] = act(sum_t sum_c A[r+t-taps/2, c] W[t][c][n] + bias[n]) (+ residual), zero halo per segment.
// Reference ops replaced: torch.nn.Conv1d(+BatchNorm1d eval, folded)+ReLU/Tanh, torch.nn.Linear,
// torch.nn.Embedding gather (encoder_sa.py:134-140, decoder_sa.py:274-286, variance_predictor.py:86-87).
#include "common.cuh"

namespace fcl {

constexpr int BM = 64, BN = 64, BK = 16, AS_LD = BM + 4;

__global__ void __launch_bounds__(256)
conv_gemm_f32_kernel(FclConvGemmParams p) {
  __shared__ __align__(16) float As[2][BK][AS_LD];
  __shared__ __align__(16) float Ws[2][BK][BN];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

  const int a_row = tid >> 2, a_k = (tid & 3) * 4;
  const int w_k = tid >> 4, w_n = (tid & 15) * 4;
  const int grow = m0 + a_row;
  const bool row_ok = grow < p.rows;
  int seg_lo = 0, seg_hi = 0x7fffffff;
  if (row_ok && p.seg_lo) { seg_lo = p.seg_lo[grow]; seg_hi = p.seg_hi[grow]; }
  const int half = p.taps >> 1;
  const int kchunks = p.cin / BK;
  const int iters = p.taps * kchunks;
  const bool wn_ok = (n0 + w_n) < p.cout;

  auto load_a = [&](int it) -> float4 {
    const int t = it / kchunks, kc = (it - t * kchunks) * BK;
    const int src = grow + t - half;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row_ok && src >= seg_lo && src < seg_hi) {
      const float* base = p.gather ? p.a + (size_t)p.gather[src] * p.lda : p.a + (size_t)src * p.lda;
      v = __ldg(reinterpret_cast<const float4*>(base + kc + a_k));
    }
    return v;
  };
  auto load_w = [&](int it) -> float4 {
    const int t = it / kchunks, kc = (it - t * kchunks) * BK;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (wn_ok) v = __ldg(reinterpret_cast<const float4*>(p.w + ((size_t)t * p.cin + kc + w_k) * p.cout + n0 + w_n));
    return v;
  };
  auto store_tiles = [&](int buf, float4 a, float4 w) {
    As[buf][a_k + 0][a_row] = a.x; As[buf][a_k + 1][a_row] = a.y;
    As[buf][a_k + 2][a_row] = a.z; As[buf][a_k + 3][a_row] = a.w;
    *reinterpret_cast<float4*>(&Ws[buf][w_k][w_n]) = w;
  };

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float4 ra = load_a(0), rw = load_w(0);
  store_tiles(0, ra, rw);
  __syncthreads();
  for (int it = 0; it < iters; ++it) {
    const int buf = it & 1;
    if (it + 1 < iters) { ra = load_a(it + 1); rw = load_w(it + 1); }
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Ws[buf][kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (it + 1 < iters) store_tiles(buf ^ 1, ra, rw);
    __syncthreads();
  }

  const int n = n0 + tx * 4;
  if (n >= p.cout) return;
  float4 bias = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p.bias) bias = __ldg(reinterpret_cast<const float4*>(p.bias + n));
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = m0 + ty * 4 + i;
    if (r >= p.rows) continue;
    float v[4] = {acc[i][0] + bias.x, acc[i][1] + bias.y, acc[i][2] + bias.z, acc[i][3] + bias.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (p.act == FCL_ACT_RELU) v[j] = fmaxf(v[j], 0.f);
      else if (p.act == FCL_ACT_TANH) v[j] = tanhf(v[j]);
    }
    if (p.residual) {
      const float4 rr = __ldg(reinterpret_cast<const float4*>(p.residual + (size_t)r * p.ldr + n));
      v[0] += rr.x; v[1] += rr.y; v[2] += rr.z; v[3] += rr.w;
    }
    *reinterpret_cast<float4*>(p.out + (size_t)r * p.ldo + n) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

}  // namespace fcl

extern "C" int fcl_conv_gemm_f32(const FclConvGemmParams* p, void* stream) {
  using namespace fcl;
  FCL_REQUIRE(p && p->a && p->w && p->out, "null pointer");
  FCL_REQUIRE(p->rows > 0 && p->cin > 0 && p->cout > 0 && p->taps >= 1 && (p->taps & 1), "bad sizes");
  FCL_REQUIRE(p->cin % BK == 0, "cin must be a multiple of 16");
  FCL_REQUIRE(p->cout % 4 == 0 && p->lda % 4 == 0 && p->ldo % 4 == 0, "cout/lda/ldo must be multiples of 4");
  FCL_REQUIRE(p->taps == 1 || (p->seg_lo && p->seg_hi), "taps > 1 needs segment bounds");
  FCL_REQUIRE(!p->residual || p->ldr % 4 == 0, "ldr must be a multiple of 4");
  dim3 grid((p->rows + BM - 1) / BM, (p->cout + BN - 1) / BN);
  conv_gemm_f32_kernel<<<grid, 256, 0, as_stream(stream)>>>(*p);
  return check_launch("fcl_conv_gemm_f32");
}
