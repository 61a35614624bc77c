// K4 (fp32 parity path): persistent decoder step loop on CUDA cores.
// Reference: nets/modules/decoder_sa.py:577-617 (loop), :146-158 (Prenet, always-on dropout),
// :63-96 (ZoneOutCell around torch.nn.LSTMCell), :398 (feat_out), fused with the ragged gather :619-630.
// One CTA owns a tile of R duration-sorted rows for all of their steps. Activations live in shared
// memory transposed ([k][row]) so a thread reads 8 rows with two LDS.128; weights stream from L2 in
// [k][n] layout (n contiguous, gates i,f,g,o of a unit adjacent) so the LSTM cell update happens in
// registers. Cell state c lives in a global scratch (L2-resident), z in shared memory.
#include "common.cuh"

namespace fcl {

constexpr int kDecThreads = 512;
constexpr int RT = 8;                         // rows per thread

template <int R> struct DecCfg {
  static constexpr int RG = R / RT;           // row groups
  static constexpr int CT = kDecThreads / RG; // column threads (each owns 4 consecutive columns)
  static constexpr int LD = R;                // smem row stride of the transposed activations
};

// acc[i][j] += sum_k A_s[k][r0+i] * W[k][col+j]
__device__ __forceinline__ void fma_panel(float (&acc)[RT][4], const float* a_s, int lda,
                                          const float* __restrict__ w, int ldw, int K) {
#pragma unroll 8
  for (int k = 0; k < K; ++k) {
    const float4 w4 = __ldg(reinterpret_cast<const float4*>(w + (size_t)k * ldw));
    const float4 a0 = *reinterpret_cast<const float4*>(a_s + k * lda);
    const float4 a1 = *reinterpret_cast<const float4*>(a_s + k * lda + 4);
    const float av[RT] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
    for (int i = 0; i < RT; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
  }
}

template <int R>
__global__ void __launch_bounds__(kDecThreads, 1)
decoder_f32_kernel(FclDecoderParams p) {
  using C = DecCfg<R>;
  constexpr int LD = C::LD;
  extern __shared__ __align__(16) float smem[];
  const int O = p.odim, U = p.prenet_units, H = p.dunits;
  float* x0_s = smem;                 // [O][R]  previous output frame
  float* x1_s = x0_s + O * LD;        // [U][R]
  float* x2_s = x1_s + U * LD;        // [U][R]
  float* z0_s = x2_s + U * LD;        // [H][R]
  float* z1_s = z0_s + H * LD;        // [H][R]
  int* meta = reinterpret_cast<int*>(z1_s + H * LD);
  int* m_row = meta;                  // original row id or -1
  int* m_dur = meta + R;
  int* m_foff = meta + 2 * R;
  int* m_utt = meta + 3 * R;
  int* m_ph = meta + 4 * R;

  const int tid = threadIdx.x;
  const int cg = tid % C::CT, rg = tid / C::CT;
  const int r0 = rg * RT;
  const int i0 = blockIdx.x * R;

  if (tid < R) {
    const int i = i0 + tid;
    int row = -1, d = 0, fo = 0, ut = 0, ph = 0;
    if (i < p.n_rows) {
      row = p.order[i];
      d = min(max(p.dur[row], 0), FCL_MAX_DURATION);
      fo = p.frame_off[row];
      ut = p.row_utt[row];
      ph = p.row_phone[row];
    }
    m_row[tid] = row; m_dur[tid] = d; m_foff[tid] = fo; m_utt[tid] = ut; m_ph[tid] = ph;
  }
  for (int i = tid; i < (O + 2 * U + 2 * H) * LD; i += kDecThreads) smem[i] = 0.f;
  __syncthreads();
  const int steps = m_dur[0];                       // rows are duration-descending
  if (steps == 0) return;

  // zero this tile's cell state
  for (int i = tid; i < R * H; i += kDecThreads) {
    const int r = i / H, u = i - r * H;
    if (m_row[r] >= 0) {
      p.cstate[(size_t)m_row[r] * H + u] = 0.f;
      p.cstate[((size_t)p.n_rows + m_row[r]) * H + u] = 0.f;
    }
  }
  __syncthreads();

  const float zo = p.zoneout, zk = 1.0f - p.zoneout;
  const bool use_drop = p.dropout_p > 0.f;
  const uint32_t drop_thr = dropout_threshold16(p.dropout_p);
  const float drop_scale = use_drop ? 1.0f / (1.0f - p.dropout_p) : 1.0f;
  const int H4 = 4 * H;
  constexpr int kMaxChunks = 4;                     // H / CT <= 4  (H <= 1024 with R=16, <= 512 with R=32)
  const int chunks = H / C::CT;

  for (int m = 0; m < steps; ++m) {
    // ---------------- prenet layer 0 and 1 ----------------
#pragma unroll 1
    for (int layer = 0; layer < 2; ++layer) {
      const float* a_s = layer == 0 ? x0_s : x1_s;
      float* o_s = layer == 0 ? x1_s : x2_s;
      const float* w = layer == 0 ? p.wp0 : p.wp1;
      const float* b = layer == 0 ? p.bp0 : p.bp1;
      const int K = layer == 0 ? O : U;
      for (int col = cg * 4; col < U; col += C::CT * 4) {
        float acc[RT][4];
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(b + col));
#pragma unroll
        for (int i = 0; i < RT; ++i) { acc[i][0] = b4.x; acc[i][1] = b4.y; acc[i][2] = b4.z; acc[i][3] = b4.w; }
        fma_panel(acc, a_s + r0, LD, w + col, U, K);
#pragma unroll
        for (int i = 0; i < RT; ++i) {
          float v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] = fmaxf(acc[i][j], 0.f);
          if (use_drop) {
            const Philox4 rnd = dropout_words(p.dropout_seed, (uint32_t)m_utt[r0 + i], (uint32_t)m_ph[r0 + i],
                                              (uint32_t)m, (uint32_t)layer, (uint32_t)(col >> 3));
            const uint32_t wa = (col & 4) ? rnd.z : rnd.x, wb = (col & 4) ? rnd.w : rnd.y;   // lanes of units col..col+3
            v[0] = (wa & 0xFFFFu) >= drop_thr ? v[0] * drop_scale : 0.f;
            v[1] = (wa >> 16) >= drop_thr ? v[1] * drop_scale : 0.f;
            v[2] = (wb & 0xFFFFu) >= drop_thr ? v[2] * drop_scale : 0.f;
            v[3] = (wb >> 16) >= drop_thr ? v[3] * drop_scale : 0.f;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) o_s[(col + j) * LD + r0 + i] = v[j];
        }
      }
      __syncthreads();
    }

    // ---------------- the two zoneout LSTM cells ----------------
#pragma unroll 1
    for (int layer = 0; layer < 2; ++layer) {
      float znew[kMaxChunks][RT];
      float* zs = layer == 0 ? z0_s : z1_s;
      float* cst = p.cstate + (size_t)layer * p.n_rows * H;
#pragma unroll
      for (int ch = 0; ch < kMaxChunks; ++ch) {
        if (ch < chunks) {
          const int u = ch * C::CT + cg;            // hidden unit; its 4 gate columns are 4u..4u+3
          float acc[RT][4];
          if (layer == 0) {
            const float4 wp = __ldg(reinterpret_cast<const float4*>(p.wpos + 4 * u));
#pragma unroll
            for (int i = 0; i < RT; ++i) {
              const int row = m_row[r0 + i];
              float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
              float pos = 0.f;
              if (row >= 0) {
                g = __ldg(reinterpret_cast<const float4*>(p.g0h + (size_t)row * H4 + 4 * u));
                const int d = m_dur[r0 + i];
                pos = m < d ? __fdiv_rn((float)m, (float)d) : 0.f;    // pad_list(..., 0) beyond d
              }
              acc[i][0] = fmaf(pos, wp.x, g.x); acc[i][1] = fmaf(pos, wp.y, g.y);
              acc[i][2] = fmaf(pos, wp.z, g.z); acc[i][3] = fmaf(pos, wp.w, g.w);
            }
            fma_panel(acc, x2_s + r0, LD, p.w0 + 4 * u, H4, U);
            fma_panel(acc, z0_s + r0, LD, p.w0 + (size_t)U * H4 + 4 * u, H4, H);
          } else {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.b1 + 4 * u));
#pragma unroll
            for (int i = 0; i < RT; ++i) { acc[i][0] = b4.x; acc[i][1] = b4.y; acc[i][2] = b4.z; acc[i][3] = b4.w; }
            fma_panel(acc, z0_s + r0, LD, p.w1 + 4 * u, H4, H);
            fma_panel(acc, z1_s + r0, LD, p.w1 + (size_t)H * H4 + 4 * u, H4, H);
          }
#pragma unroll
          for (int i = 0; i < RT; ++i) {
            const int row = m_row[r0 + i];
            const float cold = row >= 0 ? cst[(size_t)row * H + u] : 0.f;
            const float ig = sigmoid_acc(acc[i][0]), fg = sigmoid_acc(acc[i][1]);
            const float gg = tanhf(acc[i][2]), og = sigmoid_acc(acc[i][3]);
            const float cn = fg * cold + ig * gg;
            const float hn = og * tanhf(cn);
            const float zold = zs[u * LD + r0 + i];
            znew[ch][i] = zo * zold + zk * hn;                      // decoder_sa.py:95-96 (eval blend)
            if (row >= 0) cst[(size_t)row * H + u] = zo * cold + zk * cn;
          }
        }
      }
      __syncthreads();                                               // all reads of z (A operand) done
#pragma unroll
      for (int ch = 0; ch < kMaxChunks; ++ch) {
        if (ch < chunks) {
          const int u = ch * C::CT + cg;
#pragma unroll
          for (int i = 0; i < RT; ++i) zs[u * LD + r0 + i] = znew[ch][i];
        }
      }
      __syncthreads();
    }

    // ---------------- feat_out + ragged store ----------------
    for (int col = cg * 4; col < O; col += C::CT * 4) {
      float acc[RT][4];
#pragma unroll
      for (int i = 0; i < RT; ++i) {
        const int row = m_row[r0 + i];
        float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row >= 0) y = __ldg(reinterpret_cast<const float4*>(p.y0h + (size_t)row * O + col));
        acc[i][0] = y.x; acc[i][1] = y.y; acc[i][2] = y.z; acc[i][3] = y.w;
      }
      fma_panel(acc, z1_s + r0, LD, p.wf + col, O, H);
#pragma unroll
      for (int i = 0; i < RT; ++i) {
        const bool live = m_row[r0 + i] >= 0 && m < m_dur[r0 + i];
        float4 nxt = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);      // next step's prenet input
        if (p.tf_y) {                                                 // teacher forcing: prev_out = y (decoder_sa.py:510)
          nxt = make_float4(0.f, 0.f, 0.f, 0.f);
          if (live) nxt = __ldg(reinterpret_cast<const float4*>(p.tf_y + ((size_t)m_foff[r0 + i] + m) * O + col));
        }
        x0_s[(col + 0) * LD + r0 + i] = nxt.x; x0_s[(col + 1) * LD + r0 + i] = nxt.y;
        x0_s[(col + 2) * LD + r0 + i] = nxt.z; x0_s[(col + 3) * LD + r0 + i] = nxt.w;
        if (live)                                                     // exhausted rows are masked (decoder_sa.py:625-629)
          *reinterpret_cast<float4*>(p.before + ((size_t)m_foff[r0 + i] + m) * O + col) =
              make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      }
    }
    __syncthreads();
  }
}

template <int R>
static int launch_decoder(const FclDecoderParams& p, cudaStream_t s) {
  const size_t smem = (size_t)(p.odim + 2 * p.prenet_units + 2 * p.dunits) * R * sizeof(float) + 5 * R * sizeof(int);
  if (smem > 227 * 1024) { set_error("fcl_decoder_f32: tile needs %zu B shared memory", smem); return FCL_EUNSUPPORTED; }
  cudaError_t e = cudaFuncSetAttribute(decoder_f32_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("fcl_decoder_f32: %s", cudaGetErrorString(e)); return FCL_ECUDA; }
  const int tiles = (p.n_rows + R - 1) / R;
  decoder_f32_kernel<R><<<tiles, kDecThreads, smem, s>>>(p);
  return check_launch("fcl_decoder_f32");
}

}  // namespace fcl

extern "C" int fcl_decoder_f32(const FclDecoderParams* p, void* stream) {
  using namespace fcl;
  FCL_REQUIRE(p && p->order && p->dur && p->frame_off && p->row_utt && p->row_phone && p->g0h && p->y0h && p->wp0 &&
                  p->bp0 && p->wp1 && p->bp1 && p->w0 && p->wpos && p->w1 && p->b1 && p->wf && p->cstate && p->before,
              "null pointer");
  FCL_REQUIRE(p->n_rows > 0, "empty batch");
  FCL_REQUIRE(p->odim % 4 == 0 && p->prenet_units % 4 == 0, "odim/prenet_units must be multiples of 4");
  FCL_REQUIRE(p->tile_rows == 16 || p->tile_rows == 32, "tile_rows must be 16 or 32");
  FCL_REQUIRE(p->zoneout >= 0.f && p->zoneout < 1.f && p->dropout_p >= 0.f && p->dropout_p < 1.f, "bad rates");
  const int ct = kDecThreads / (p->tile_rows / RT);
  FCL_REQUIRE(p->dunits % ct == 0 && p->dunits / ct <= 4, "dunits must be a multiple of the column-thread count (<= 4 chunks)");
  cudaStream_t s = as_stream(stream);
  return p->tile_rows == 16 ? launch_decoder<16>(*p, s) : launch_decoder<32>(*p, s);
}
