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
    const int r = tid;
    const int grow = m0 + r;
    const bool row_ok = grow < p.rows;
    int seg_lo = 0, seg_hi = 0x7fffffff;
    if (row_ok && p.seg_lo) { seg_lo = p.seg_lo[grow]; seg_hi = p.seg_hi[grow]; }
    const int half = p.taps >> 1;
    for (int it = 0; it < iters; ++it) {
      const int s = it % stages;
      const uint32_t ph = (uint32_t)(it / stages) & 1u;
      const int t = it / kchunks, kc = (it - t * kchunks) * kstage;
      const int src = grow + t - half;
      const bool valid = row_ok && src >= seg_lo && src < seg_hi;
      const float* base = nullptr;
      if (valid) {
        const size_t arow = p.gather ? (size_t)p.gather[src] : p.row_gather ? (size_t)p.row_gather[src] : (size_t)src;
        base = p.a + arow * p.lda + kc;
      }
      // issue all global loads of the stage before touching shared memory (memory-level parallelism)
      float4 v[20];                                   // kstage <= 80 -> 20 float4
#pragma unroll
      for (int j = 0; j < 20; ++j) {
        v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j * 4 < kstage && valid) v[j] = __ldg(reinterpret_cast<const float4*>(base) + j);
      }
      mbar_wait(&empty_bar[s], ph ^ 1u);
      uint8_t* a_s = smem + (size_t)s * stage_bytes + (size_t)r * 16;
#pragma unroll
      for (int j = 0; j < 10; ++j) {
        if (j * 8 < kstage) {
          uint4 w;
          w.x = pack_bf16(v[2 * j].x, v[2 * j].y); w.y = pack_bf16(v[2 * j].z, v[2 * j].w);
          w.z = pack_bf16(v[2 * j + 1].x, v[2 * j + 1].y); w.w = pack_bf16(v[2 * j + 1].z, v[2 * j + 1].w);
          *reinterpret_cast<uint4*>(a_s + (size_t)j * 2048) = w;
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&full_bar[s]);
    }
    // ------------------------------------------------ epilogue
    mbar_wait(&accum_bar, 0);
    tc_fence_after();
    const int n0 = nt * ntile;
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    for (int g = 0; g < ntile / 16; ++g) {
      float acc[16];
      tmem_ld16(lane_addr + (uint32_t)(g * 16), acc);
      if (row_ok || p.out_layout == 1) {
        const int n = n0 + g * 16;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 o = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
          if (p.bias) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n) + q);
            o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
          }
          if (p.act == FCL_ACT_RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
          else if (p.act == FCL_ACT_TANH) { o.x = tanhf(o.x); o.y = tanhf(o.y); o.z = tanhf(o.z); o.w = tanhf(o.w); }
          if (p.residual && row_ok) {
            const float4 rr = __ldg(reinterpret_cast<const float4*>(p.residual + (size_t)grow * p.ldr + n) + q);
            o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
          }
          if (p.out_layout == 0)
            reinterpret_cast<float4*>(p.out + (size_t)grow * p.ldo + n)[q] = o;
          else   // tile-transposed [row tile][cout/4][128][4]: a warp stores 512 contiguous bytes
            *reinterpret_cast<float4*>(p.out + (((size_t)blockIdx.x * (p.cout >> 2) + ((n >> 2) + q)) * 128 + r) * 4) = o;
        }
      }
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
  const size_t smem = stage_bytes * stages;
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
