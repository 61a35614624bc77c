// Ragged conv1d / linear as an implicit GEMM on the 5th-gen tensor cores (tcgen05, bf16 x bf16 -> fp32 in TMEM).
// Same contract as fcl_conv_gemm_f32 (zero halo per segment, optional embedding gather, bias / activation /
// residual epilogue, fp32 activations in HBM) -- operands are rounded to bf16, accumulation is fp32.
// Reference ops replaced: torch.nn.Conv1d(+BatchNorm1d eval, folded)+ReLU/Tanh, torch.nn.Linear,
// torch.nn.Embedding (encoder_sa.py:134-140, decoder_sa.py:274-286, variance_predictor.py:86-87).
//
// CTA = one 128-row x ntile-column output tile. Warp roles:
//   warps 0-3  A producers: thread r loads row (r + tap - taps/2) of the fp32 activations (masked at the
//              utterance boundary), converts to bf16 and writes the UMMA core-matrix image into the stage;
//              afterwards the same warps run the epilogue (TMEM -> registers -> bias/act/residual -> HBM)
//   warp 4     B producer: one cp.async.bulk per stage from the pre-packed bf16 weights (L2-resident)
//   warp 5     MMA issuer: one thread issues tcgen05.mma (M=128, N=ntile, K=16) x kstage/16 per stage
// Stages are handed over with mbarriers (full: 128 thread arrivals + bulk-copy bytes; empty: tcgen05.commit).
#include "common.cuh"
#include "umma.cuh"

namespace fcl {
using namespace umma;

constexpr int kGemmThreads = 192;
constexpr int kMaxStages = 4;

__global__ void __launch_bounds__(kGemmThreads)
conv_gemm_bf16_kernel(FclConvGemmBf16Params p, int stages) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], accum_bar;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int m0 = blockIdx.x * 128;
  const int nt = blockIdx.y;
  const int ntile = p.ntile, kstage = p.kstage;
  const uint32_t a_bytes = 128u * kstage * 2u, b_bytes = (uint32_t)ntile * kstage * 2u;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const int kchunks = p.cin / kstage;
  const int iters = p.taps * kchunks;
  const uint32_t tmem_cols = tmem_cols_pow2((uint32_t)ntile);

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[s], 128 + 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(&tmem_base_s, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp < 4) {
    // ------------------------------------------------ A producer
    // Coalesced loads: one warp instruction covers 8 rows x 64 contiguous bytes (lane & 7 -> row, lane >> 3 ->
    // 16-byte column), so every 32-byte sector fetched is fully used; a thread then owns 4 consecutive k of a
    // row and stores them as 8 bytes of that row's 16-byte core-matrix line (conflict-free within 16 lanes).
    const int lane = tid & 31;
    const int kq = lane >> 3;                           // which float4 of a 16-float group
    const int half = p.taps >> 1;
    int rloc[4], grow4[4], lo4[4], hi4[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      rloc[g] = warp * 32 + g * 8 + (lane & 7);
      grow4[g] = m0 + rloc[g];
      lo4[g] = 0; hi4[g] = grow4[g] < p.rows ? 0x7fffffff : 0;           // hi = 0 masks rows beyond the matrix
      if (grow4[g] < p.rows && p.seg_lo) { lo4[g] = p.seg_lo[grow4[g]]; hi4[g] = p.seg_hi[grow4[g]]; }
    }
    const int kgroups = kstage >> 4;                    // 16-float groups per stage (<= 5)
    for (int it = 0; it < iters; ++it) {
      const int s = it % stages;
      const uint32_t ph = (uint32_t)(it / stages) & 1u;
      const int t = it / kchunks, kc = (it - t * kchunks) * kstage;
      float4 v[4][5];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int src = grow4[g] + t - half;
        const bool valid = src >= lo4[g] && src < hi4[g];
        const float* base = nullptr;
        if (valid) {
          const size_t arow = p.gather ? (size_t)p.gather[src] : p.row_gather ? (size_t)p.row_gather[src] : (size_t)src;
          base = p.a + arow * p.lda + kc + 4 * kq;
        }
#pragma unroll
        for (int kg = 0; kg < 5; ++kg) {
          v[g][kg] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (kg < kgroups && valid) v[g][kg] = __ldg(reinterpret_cast<const float4*>(base + 16 * kg));
        }
      }
      mbar_wait(&empty_bar[s], ph ^ 1u);
      uint8_t* a_s = smem + (size_t)s * stage_bytes;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
#pragma unroll
        for (int kg = 0; kg < 5; ++kg) {
          if (kg < kgroups) {
            const int slab = 2 * kg + (kq >> 1);        // k / 8 with k = 16 kg + 4 kq
            *reinterpret_cast<uint2*>(a_s + (size_t)slab * 2048 + (size_t)rloc[g] * 16 + (kq & 1) * 8) =
                make_uint2(pack_bf16(v[g][kg].x, v[g][kg].y), pack_bf16(v[g][kg].z, v[g][kg].w));
          }
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&full_bar[s]);
    }
    // ------------------------------------------------ epilogue
    // TMEM -> registers (thread = row) -> bias/activation -> this warp's shared-memory patch (the pipeline
    // stages are idle by now) -> coalesced rows (16 lanes x 16 B per row) + residual -> HBM.
    mbar_wait(&accum_bar, 0);
    tc_fence_after();
    const int n0 = nt * ntile;
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    constexpr int kPatchLd = 68;                                          // floats per patch row (64 + pad)
    float* patch = reinterpret_cast<float*>(smem) + (size_t)warp * 32 * kPatchLd;
    for (int c0 = 0; c0 < ntile; c0 += 64) {
      const int sw = min(64, ntile - c0);
      for (int g = 0; g < sw / 16; ++g) {
        float acc[16];
        tmem_ld16(lane_addr + (uint32_t)(c0 + g * 16), acc);
        const int n = n0 + c0 + g * 16;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 o = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
          if (p.bias) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n) + q);
            o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
          }
          if (p.act == FCL_ACT_RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
          else if (p.act == FCL_ACT_TANH) { o.x = tanhf(o.x); o.y = tanhf(o.y); o.z = tanhf(o.z); o.w = tanhf(o.w); }
          *reinterpret_cast<float4*>(patch + lane * kPatchLd + g * 16 + q * 4) = o;
        }
      }
      __syncwarp();
      const int col = (lane & 15) * 4;
      if (col < sw) {
#pragma unroll 4
        for (int rr = 0; rr < 32; rr += 2) {
          const int rl = rr + (lane >> 4);
          const int grow = m0 + warp * 32 + rl;
          if (grow < p.rows) {
            float4 o = *reinterpret_cast<const float4*>(patch + rl * kPatchLd + col);
            if (p.residual) {
              const float4 r4 = __ldg(reinterpret_cast<const float4*>(p.residual + (size_t)grow * p.ldr + n0 + c0 + col));
              o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;
            }
            if (p.out_bf16)
              *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)grow * p.ldo + n0 + c0 + col) =
                  make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
            else
              *reinterpret_cast<float4*>(p.out + (size_t)grow * p.ldo + n0 + c0 + col) = o;
          }
        }
      }
      __syncwarp();
    }
  } else if (warp == 4) {
    // ------------------------------------------------ B producer
    if (elect_one()) {
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.w_packed) + (size_t)nt * iters * b_bytes;
      for (int it = 0; it < iters; ++it) {
        const int s = it % stages;
        const uint32_t ph = (uint32_t)(it / stages) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        mbar_arrive_expect_tx(&full_bar[s], b_bytes);
        bulk_g2s(smem + (size_t)s * stage_bytes + a_bytes, wsrc + (size_t)it * b_bytes, b_bytes, &full_bar[s]);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------ MMA issuer
    if (elect_one()) {
      const uint32_t idesc = idesc_bf16_f32(128u, (uint32_t)ntile);
      const uint32_t b_lbo = (uint32_t)ntile * 16u;
      for (int it = 0; it < iters; ++it) {
        const int s = it % stages;
        const uint32_t ph = (uint32_t)(it / stages) & 1u;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + (size_t)s * stage_bytes);
        const uint32_t b_addr = a_addr + a_bytes;
        for (int k = 0; k < kstage / 16; ++k) {
          const uint64_t ad = smem_desc(a_addr + (uint32_t)k * 2u * 2048u, 2048u, 128u);
          const uint64_t bd = smem_desc(b_addr + (uint32_t)k * 2u * b_lbo, b_lbo, 128u);
          mma_bf16_ss(tmem, ad, bd, idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        mma_commit(&empty_bar[s]);          // frees the stage when these MMAs retire
      }
      mma_commit(&accum_bar);               // accumulator complete
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, tmem_cols);
}

}  // namespace fcl

extern "C" int fcl_conv_gemm_bf16(const FclConvGemmBf16Params* p, void* stream) {
  using namespace fcl;
  FCL_REQUIRE(p && p->a && p->w_packed && p->out, "null pointer");
  FCL_REQUIRE(p->rows > 0 && p->cin > 0 && p->cout > 0 && p->taps >= 1 && (p->taps & 1), "bad sizes");
  FCL_REQUIRE(p->kstage % 16 == 0 && p->kstage >= 16 && p->kstage <= 80 && p->cin % p->kstage == 0,
              "kstage must be a multiple of 16 (<= 80) dividing cin");
  FCL_REQUIRE(p->ntile % 16 == 0 && p->ntile >= 16 && p->ntile <= 256 && p->cout % p->ntile == 0,
              "ntile must be a multiple of 16 (<= 256) dividing cout");
  FCL_REQUIRE(p->lda % 4 == 0 && p->ldo % 4 == 0 && (!p->residual || p->ldr % 4 == 0), "leading dims must be multiples of 4");
  FCL_REQUIRE(p->taps == 1 || (p->seg_lo && p->seg_hi), "taps > 1 needs segment bounds");
  const size_t stage_bytes = (size_t)(128 + p->ntile) * p->kstage * 2;
  // prefer a footprint that lets two CTAs share an SM (one CTA's epilogue overlaps the other's main loop)
  int stages = (int)((108 * 1024) / stage_bytes);
  if (stages < 2) stages = (int)((216 * 1024) / stage_bytes);
  stages = stages > kMaxStages ? kMaxStages : stages;
  FCL_REQUIRE(stages >= 2, "tile too large for two pipeline stages");
  size_t smem = stage_bytes * stages;
  if (smem < 4 * 32 * 68 * sizeof(float)) smem = 4 * 32 * 68 * sizeof(float);   // epilogue staging patches
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024);
    if (e != cudaSuccess) { set_error("fcl_conv_gemm_bf16: %s", cudaGetErrorString(e)); return FCL_ECUDA; }
    attr_done = true;
  }
  dim3 grid((p->rows + 127) / 128, p->cout / p->ntile);
  conv_gemm_bf16_kernel<<<grid, kGemmThreads, smem, as_stream(stream)>>>(*p, stages);
  return check_launch("fcl_conv_gemm_bf16");
}
