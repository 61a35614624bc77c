// Ragged conv1d / linear as an implicit GEMM on the 5th-gen tensor cores (tcgen05, bf16 x bf16 -> fp32 in TMEM).
// Same contract as fcl_conv_gemm_f32 (zero halo per utterance, optional embedding gather, bias / activation /
// residual epilogue, fp32 activations in HBM) -- operands are rounded to bf16, accumulation is fp32.
// Reference ops replaced: torch.nn.Conv1d(+BatchNorm1d eval, folded)+ReLU/Tanh, torch.nn.Linear,
// torch.nn.Embedding (encoder_sa.py:134-140, decoder_sa.py:274-286, variance_predictor.py:86-87).
//
// CTA = one tile of up to 128 output rows x ntile columns. For a k-tap convolution the tile never crosses an
// utterance: its (128 + 2*halo)-row input WINDOW (rows outside the utterance are zero) is loaded from HBM
// exactly once per 64-channel K stage, converted to bf16 and laid out as a UMMA operand image in shared
// memory; the taps are then just descriptor offsets: tap t multiplies image rows [t, t+128) -- a start
// address shifted by t * 16 bytes -- so no input row is fetched twice. Which global row feeds which window
// row (and where each output row goes) comes from per-tile maps built by fcl_conv_tiles.
// Warp roles:
//   warps 0-3  A producers: coalesced fp32 loads -> bf16 image stage; afterwards the same warps run the
//              epilogue (TMEM -> registers -> bias/act -> shared-memory patch -> coalesced rows (+residual) -> HBM)
//   warp 4     B producer: one cp.async.bulk per (K stage, tap) from the pre-packed bf16 weights (L2-resident)
//   warp 5     MMA issuer: one thread issues tcgen05.mma (M=128, N=ntile, K=16)
// Stages are handed over with mbarriers (A full: 128 thread arrivals; B full: bulk-copy bytes; empty: tcgen05.commit).
#include "common.cuh"
#include "umma.cuh"

namespace fcl {
using namespace umma;

constexpr int kGemmThreads = 192;
constexpr int kWinRows = 136;                       // image rows per slab (128 + 2*halo <= 132, padded to 8)
constexpr uint32_t kSlabBytes = kWinRows * 16;      // 2176: LBO of the A image
constexpr int kMaxAStages = 3, kMaxBStages = 6;

struct GemmShared {
  uint64_t a_full[kMaxAStages], a_empty[kMaxAStages];
  uint64_t b_full[kMaxBStages], b_empty[kMaxBStages];
  uint64_t accum;
  uint32_t tmem_base;
  int src[kWinRows];                                // global source row of each window row, -1 = zero
  int dst[128];                                     // global destination row of each output row, -1 = none
};

__global__ void __launch_bounds__(kGemmThreads)
conv_gemm_bf16_kernel(FclConvGemmBf16Params p, int a_stages, int b_stages) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ GemmShared sh;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x, nt = blockIdx.y;
  if (p.n_tiles_dev && tile >= *p.n_tiles_dev) return;          // grid is an upper bound for data-dependent tilings
  const int ntile = p.ntile, kstage = p.kstage, taps = p.taps, halo = taps >> 1;
  const int map_halo = p.tile_src ? p.map_halo : halo;   // without maps the window is the tile itself (taps == 1)
  const int win = 128 + 2 * map_halo;
  const int slabs = kstage >> 3;
  const uint32_t a_bytes = (uint32_t)slabs * kSlabBytes, b_bytes = (uint32_t)ntile * kstage * 2u;
  const int kchunks = p.cin / kstage;
  const uint32_t tmem_cols = tmem_cols_pow2((uint32_t)ntile);
  uint8_t* a_ring = smem;
  uint8_t* b_ring = smem + (size_t)a_stages * a_bytes;

  // ---- tile maps
  for (int j = tid; j < kWinRows; j += kGemmThreads) {
    int s = -1;
    if (j < win) {
      if (p.tile_src) s = p.tile_src[(size_t)tile * kWinRows + j];
      else { const int g = tile * 128 + j; s = g < p.rows ? g : -1; }            // taps == 1: identity
    }
    sh.src[j] = s;
  }
  for (int j = tid; j < 128; j += kGemmThreads) {
    int d;
    if (p.tile_dst) d = p.tile_dst[(size_t)tile * 128 + j];
    else { const int g = tile * 128 + j; d = g < p.rows ? g : -1; }
    sh.dst[j] = d;
  }
  if (tid == 0) {
    for (int s = 0; s < a_stages; ++s) { mbar_init(&sh.a_full[s], 128); mbar_init(&sh.a_empty[s], 1); }
    for (int s = 0; s < b_stages; ++s) { mbar_init(&sh.b_full[s], 1); mbar_init(&sh.b_empty[s], 1); }
    mbar_init(&sh.accum, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(&sh.tmem_base, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh.tmem_base;

  if (warp < 4) {
    // ------------------------------------------------ A producer
    // One warp instruction covers 8 window rows x 64 contiguous bytes (lane & 7 -> row, lane >> 3 -> 16-byte
    // column): every fetched sector is fully used; the thread then owns 4 consecutive k of a row and stores
    // them as 8 bytes of that row's 16-byte core-matrix line.
    const int kq = lane >> 3;
    const int kgroups = kstage >> 4;                  // 16-float groups per stage (<= 5)
    constexpr int kRowGroups = kWinRows / 8;          // 17 groups of 8 rows, dealt round-robin to the 4 warps
    const float* rowp[5];
    int jrow[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const int rg = warp + 4 * i;
      jrow[i] = rg * 8 + (lane & 7);
      rowp[i] = nullptr;
      if (rg < kRowGroups) {
        const int s = sh.src[jrow[i]];
        if (s >= 0) {
          const size_t arow = p.gather ? (size_t)p.gather[s] : p.row_gather ? (size_t)p.row_gather[s] : (size_t)s;
          rowp[i] = p.a + arow * p.lda + 4 * kq;
        }
      }
    }
    for (int kc = 0; kc < kchunks; ++kc) {
      const int s = kc % a_stages;
      const uint32_t ph = (uint32_t)(kc / a_stages) & 1u;
      float4 v[5][5];
#pragma unroll
      for (int i = 0; i < 5; ++i) {
#pragma unroll
        for (int kg = 0; kg < 5; ++kg) {
          v[i][kg] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (kg < kgroups && rowp[i]) v[i][kg] = __ldg(reinterpret_cast<const float4*>(rowp[i] + kc * kstage + 16 * kg));
        }
      }
      mbar_wait(&sh.a_empty[s], ph ^ 1u);
      uint8_t* a_s = a_ring + (size_t)s * a_bytes;
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        if (warp + 4 * i < kRowGroups) {
#pragma unroll
          for (int kg = 0; kg < 5; ++kg) {
            if (kg < kgroups) {
              const int slab = 2 * kg + (kq >> 1);      // k / 8 with k = 16 kg + 4 kq
              *reinterpret_cast<uint2*>(a_s + (size_t)slab * kSlabBytes + (size_t)jrow[i] * 16 + (kq & 1) * 8) =
                  make_uint2(pack_op(v[i][kg].x, v[i][kg].y), pack_op(v[i][kg].z, v[i][kg].w));
            }
          }
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&sh.a_full[s]);
    }
    // ------------------------------------------------ epilogue
    mbar_wait(&sh.accum, 0);
    tc_fence_after();
    const int n0 = nt * ntile;
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    constexpr int kPatchLd = 68;                                          // floats per patch row (64 + pad)
    float* patch = reinterpret_cast<float*>(smem) + (size_t)warp * 32 * kPatchLd;   // rings are idle by now
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
          else if (p.act == FCL_ACT_TANH) { o.x = tanh_fast(o.x); o.y = tanh_fast(o.y); o.z = tanh_fast(o.z); o.w = tanh_fast(o.w); }
          *reinterpret_cast<float4*>(patch + lane * kPatchLd + g * 16 + q * 4) = o;
        }
      }
      __syncwarp();
      const int col = (lane & 15) * 4;
      if (col < sw) {
#pragma unroll 4
        for (int rr = 0; rr < 32; rr += 2) {
          const int rl = rr + (lane >> 4);
          const int grow = sh.dst[warp * 32 + rl];
          if (grow >= 0) {
            float4 o = *reinterpret_cast<const float4*>(patch + rl * kPatchLd + col);
            if (p.residual) {
              const float4 r4 = __ldg(reinterpret_cast<const float4*>(p.residual + (size_t)grow * p.ldr + n0 + c0 + col));
              o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;
            }
            if (p.out_bf16)
              *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.out) + (size_t)grow * p.ldo + n0 + c0 + col) =
                  make_uint2(pack_op(o.x, o.y), pack_op(o.z, o.w));
            else
              *reinterpret_cast<float4*>(p.out + (size_t)grow * p.ldo + n0 + c0 + col) = o;
          }
        }
      }
      __syncwarp();
    }
  } else if (warp == 4) {
    // ------------------------------------------------ B producer: blocks ordered [column tile][K stage][tap]
    if (elect_one()) {
      const int iters = kchunks * taps;
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.w_packed) + (size_t)nt * iters * b_bytes;
      for (int it = 0; it < iters; ++it) {
        const int s = it % b_stages;
        const uint32_t ph = (uint32_t)(it / b_stages) & 1u;
        mbar_wait(&sh.b_empty[s], ph ^ 1u);
        mbar_arrive_expect_tx(&sh.b_full[s], b_bytes);
        bulk_g2s(b_ring + (size_t)s * b_bytes, wsrc + (size_t)it * b_bytes, b_bytes, &sh.b_full[s]);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------ MMA issuer
    if (elect_one()) {
      const uint32_t idesc = idesc_op_f32(128u, (uint32_t)ntile);
      const uint32_t b_lbo = (uint32_t)ntile * 16u;
      int it = 0;
      for (int kc = 0; kc < kchunks; ++kc) {
        const int sa = kc % a_stages;
        mbar_wait(&sh.a_full[sa], (uint32_t)(kc / a_stages) & 1u);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(a_ring + (size_t)sa * a_bytes);
        for (int t = 0; t < taps; ++t, ++it) {
          const int sb = it % b_stages;
          mbar_wait(&sh.b_full[sb], (uint32_t)(it / b_stages) & 1u);
          tc_fence_after();
          const uint32_t b_addr = smem_u32(b_ring + (size_t)sb * b_bytes);
          for (int k = 0; k < kstage / 16; ++k) {
            // tap t reads window rows [t, t + 128): the image start shifted by t rows of 16 bytes
            const uint64_t ad = smem_desc(a_addr + (uint32_t)(t + map_halo - halo) * 16u + (uint32_t)k * 2u * kSlabBytes,
                                          kSlabBytes, 128u);
            const uint64_t bd = smem_desc(b_addr + (uint32_t)k * 2u * b_lbo, b_lbo, 128u);
            mma_bf16_ss(tmem, ad, bd, idesc, (it > 0 || k > 0) ? 1u : 0u);
          }
          mma_commit(&sh.b_empty[sb]);        // frees the weight stage when these MMAs retire
        }
        mma_commit(&sh.a_empty[sa]);          // frees the window stage
      }
      mma_commit(&sh.accum);                  // accumulator complete
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, tmem_cols);
}

// ---------------------------------------------------------------- tile maps for k-tap convolutions
// pass 1 (one CTA): tiles per segment -> exclusive scan -> first tile of each segment, total
__global__ void __launch_bounds__(1024, 1)
conv_tiles_scan_kernel(FclConvTilesParams p) {
  __shared__ int warp_sums[32];
  __shared__ int carry;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < p.n_segs; base += 1024) {
    const int s = base + tid;
    const int n = s < p.n_segs ? (p.seg_off[s + 1] - p.seg_off[s] + 127) / 128 : 0;
    int incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      int w = warp_sums[lane], wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += y; }
      warp_sums[lane] = wi - w;
    }
    __syncthreads();
    const int excl = carry + warp_sums[wid] + incl - n;
    if (s < p.n_segs) p.seg_first_tile[s] = excl;
    __syncthreads();
    if (tid == 1023) carry = excl + n;
    __syncthreads();
  }
  if (tid == 0) { p.seg_first_tile[p.n_segs] = carry; *p.n_tiles = carry; }
}

// pass 2: one CTA per tile slot fills its window / destination maps
__global__ void __launch_bounds__(kWinRows)
conv_tiles_fill_kernel(FclConvTilesParams p) {
  const int tile = blockIdx.x;
  const int total = p.seg_first_tile[p.n_segs];
  if (tile >= total) return;
  int lo = 0, hi = p.n_segs;                       // last segment whose first tile <= tile
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (p.seg_first_tile[mid] <= tile) lo = mid; else hi = mid; }
  const int seg_lo = p.seg_off[lo], seg_hi = p.seg_off[lo + 1];
  const int start = seg_lo + (tile - p.seg_first_tile[lo]) * 128;
  const int j = threadIdx.x;
  const int src = start - p.halo + j;
  p.tile_src[(size_t)tile * kWinRows + j] = (j < 128 + 2 * p.halo && src >= seg_lo && src < seg_hi) ? src : -1;
  if (j < 128) p.tile_dst[(size_t)tile * 128 + j] = (start + j < seg_hi) ? start + j : -1;
}

}  // namespace fcl

extern "C" int fcl_conv_tiles(const FclConvTilesParams* p, void* stream) {
  using namespace fcl;
  FCL_REQUIRE(p && p->seg_off && p->seg_first_tile && p->tile_src && p->tile_dst && p->n_tiles, "null pointer");
  FCL_REQUIRE(p->n_segs > 0 && p->max_tiles > 0 && p->halo >= 0 && p->halo <= 4, "bad sizes");
  conv_tiles_scan_kernel<<<1, 1024, 0, as_stream(stream)>>>(*p);
  conv_tiles_fill_kernel<<<p->max_tiles, kWinRows, 0, as_stream(stream)>>>(*p);
  return check_launch("fcl_conv_tiles");
}

extern "C" int fcl_conv_gemm_bf16(const FclConvGemmBf16Params* p, void* stream) {
  using namespace fcl;
  FCL_REQUIRE(p && p->a && p->w_packed && p->out, "null pointer");
  FCL_REQUIRE(p->rows > 0 && p->cin > 0 && p->cout > 0 && p->taps >= 1 && (p->taps & 1) && p->taps <= 9, "bad sizes");
  FCL_REQUIRE(p->kstage % 16 == 0 && p->kstage >= 16 && p->kstage <= 80 && p->cin % p->kstage == 0,
              "kstage must be a multiple of 16 (<= 80) dividing cin");
  FCL_REQUIRE(p->ntile % 16 == 0 && p->ntile >= 16 && p->ntile <= 256 && p->cout % p->ntile == 0,
              "ntile must be a multiple of 16 (<= 256) dividing cout");
  FCL_REQUIRE(p->lda % 4 == 0 && p->ldo % 4 == 0 && (!p->residual || p->ldr % 4 == 0), "leading dims must be multiples of 4");
  FCL_REQUIRE(p->taps == 1 || (p->tile_src && p->tile_dst && p->n_tiles > 0), "taps > 1 needs tile maps (fcl_conv_tiles)");
  FCL_REQUIRE(!p->tile_src || (p->tile_dst && p->map_halo >= p->taps / 2 && p->map_halo <= 4), "tile maps need map_halo >= taps/2");
  const size_t a_bytes = (size_t)(p->kstage / 8) * kSlabBytes, b_bytes = (size_t)p->ntile * p->kstage * 2;
  // prefer a footprint that lets two CTAs share an SM (one CTA's epilogue overlaps the other's main loop)
  const int kchunks = p->cin / p->kstage;
  int a_stages = kchunks < 2 ? 1 : 2, b_stages = 0;
  size_t budget = 108 * 1024;
  for (int attempt = 0; attempt < 2 && b_stages < 2; ++attempt) {
    b_stages = (int)((budget - a_stages * a_bytes) / b_bytes);
    budget = 216 * 1024;
  }
  if (b_stages > kMaxBStages) b_stages = kMaxBStages;
  FCL_REQUIRE(b_stages >= 2, "tile too large for two weight stages");
  size_t smem = a_stages * a_bytes + b_stages * b_bytes;
  if (smem < 4 * 32 * 68 * sizeof(float)) smem = 4 * 32 * 68 * sizeof(float);   // epilogue staging patches
  if (int rc = ensure_dyn_smem(conv_gemm_bf16_kernel, 216 * 1024, "fcl_conv_gemm_bf16")) return rc;
  const int tiles = p->tile_src ? p->n_tiles : (p->rows + 127) / 128;
  dim3 grid(tiles, p->cout / p->ntile);
  conv_gemm_bf16_kernel<<<grid, kGemmThreads, smem, as_stream(stream)>>>(*p, a_stages, b_stages);
  return check_launch("fcl_conv_gemm_bf16");
}
