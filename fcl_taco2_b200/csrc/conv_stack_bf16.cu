// Fused stack of k-tap conv1d layers (postnet / encoder convs) on tcgen05: activations never leave the SM.
// Reference: nets/modules/decoder_sa.py:274-286,632 (Postnet: 5 x [Conv1d k5 -> BatchNorm eval (folded) -> tanh],
// last layer without tanh, + residual) and nets/modules/encoder_sa.py:135-140 (3 x [conv k5 -> BN -> ReLU]).
//
// CTA = one tile of S = 128 - 2*h*(L-1) output rows of one utterance (h = taps/2, L layers). Every layer's
// activations live in shared memory as a bf16 UMMA operand image [C/8][136 rows][8] whose row j is global row
// (row0 - L*h + j); a layer computes image rows [h, h+128) of the next image from rows [0, 128+2h) of the
// current one -- tap t is a 16-byte-per-row shift of the A descriptor -- so the valid region shrinks by h rows
// per side per layer (halo recompute, 12.5 % for the postnet) and only the first input and the last output
// touch HBM. Rows outside the utterance are forced to zero in every layer ("zero halo per utterance", never
// pad-and-convolve). Weights (bf16, pre-tiled) stream from L2 per (K stage, tap) through a bulk-copy ring.
// Warps 0-3: layer-0 loader, then per-layer epilogue (TMEM -> bias/act -> next image, or -> HBM (+residual));
// warp 4: weight producer; warp 5: MMA issuer. Two CTAs share an SM so one tile's epilogues overlap the
// other's MMAs.
#include "common.cuh"
#include "umma.cuh"

namespace fcl {
using namespace umma;

constexpr int kCsThreads = 192;
constexpr int kCsWinRows = 136;
constexpr uint32_t kCsSlab = kCsWinRows * 16;       // 2176 B per 8-channel slab
constexpr int kCsMaxBStages = 8;

struct CsShared {
  uint64_t img_ready[FCL_MAX_STACK_LAYERS];         // image of layer l complete (128 thread arrivals)
  uint64_t b_full[kCsMaxBStages], b_empty[kCsMaxBStages];
  uint64_t accum;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kCsThreads)
conv_stack_bf16_kernel(FclConvStackParams p, uint32_t img_bytes, uint32_t b_slot_bytes, uint32_t tmem_cols, int kCsBStages) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ CsShared sh;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x;
  if (p.n_tiles_dev && tile >= *p.n_tiles_dev) return;
  const int L = p.n_layers, taps = p.taps, h = taps >> 1;
  const int row0 = p.tiles[4 * tile], seg_lo = p.tiles[4 * tile + 1], seg_hi = p.tiles[4 * tile + 2];
  const int gbase = row0 - L * h;                    // global row of image row 0
  // ONE image buffer: a layer's epilogue runs after all of its MMAs retired, so it overwrites its own input
  uint8_t* img[2] = {smem, smem};
  uint8_t* b_ring = smem + (size_t)img_bytes;

  if (tid == 0) {
    for (int l = 0; l < L; ++l) mbar_init(&sh.img_ready[l], 128);
    for (int s = 0; s < kCsBStages; ++s) { mbar_init(&sh.b_full[s], 1); mbar_init(&sh.b_empty[s], 1); }
    mbar_init(&sh.accum, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(&sh.tmem_base, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh.tmem_base;

  if (warp < 4) {
    // ------------------------------------------------ layer-0 image: coalesced fp32 loads -> bf16
    {
      const int cin = p.layers[0].cin;               // may be zero-padded beyond the real input width
      const int cin_real = p.in_channels;
      const int kq = lane >> 3;
      for (int rg = warp; rg < kCsWinRows / 8; rg += 4) {
        const int j = rg * 8 + (lane & 7);
        const int g = gbase + j;
        const float* src = nullptr;
        if (j < 128 + 2 * h && g >= seg_lo && g < seg_hi)
          src = p.in + (p.gather ? (size_t)p.gather[g] : (size_t)g) * p.ld_in + 4 * kq;
        for (int k0 = 0; k0 < cin; k0 += 64) {         // up to four 16-float groups per pass
          float4 v[4];
#pragma unroll
          for (int kg = 0; kg < 4; ++kg) {
            v[kg] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (src && k0 + 16 * kg < cin_real) v[kg] = __ldg(reinterpret_cast<const float4*>(src + k0 + 16 * kg));
          }
#pragma unroll
          for (int kg = 0; kg < 4; ++kg) {
            if (k0 + 16 * kg < cin) {
              const int slab = (k0 >> 3) + 2 * kg + (kq >> 1);
              *reinterpret_cast<uint2*>(img[0] + (size_t)slab * kCsSlab + (size_t)j * 16 + (kq & 1) * 8) =
                  make_uint2(pack_op(v[kg].x, v[kg].y), pack_op(v[kg].z, v[kg].w));
            }
          }
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&sh.img_ready[0]);
    }
    // ------------------------------------------------ per-layer epilogue; thread = TMEM lane i = image row i + h
    const int i = tid;
    const int g = gbase + i + h;
    const bool inside = g >= seg_lo && g < seg_hi;
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    for (int l = 0; l < L; ++l) {
      const FclConvLayer& ly = p.layers[l];
      mbar_wait(&sh.accum, (uint32_t)l & 1u);
      tc_fence_after();
      if (l + 1 < L) {
        uint8_t* nxt = img[(l + 1) & 1] + (size_t)(i + h) * 16;
        for (int c0 = 0; c0 < ly.cout; c0 += 16) {
          float v[16];
          tmem_ld16(lane_addr + (uint32_t)c0, v);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(ly.bias + c0) + q);
            v[4 * q] += b.x; v[4 * q + 1] += b.y; v[4 * q + 2] += b.z; v[4 * q + 3] += b.w;
          }
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            float x = v[e];
            if (ly.act == FCL_ACT_RELU) x = fmaxf(x, 0.f);
            else if (ly.act == FCL_ACT_TANH) x = tanh_fast(x);
            v[e] = inside ? x : 0.f;                                  // zero halo at the utterance boundary
          }
#pragma unroll
          for (int k8 = 0; k8 < 2; ++k8) {
            uint4 w;
            w.x = pack_op(v[8 * k8], v[8 * k8 + 1]); w.y = pack_op(v[8 * k8 + 2], v[8 * k8 + 3]);
            w.z = pack_op(v[8 * k8 + 4], v[8 * k8 + 5]); w.w = pack_op(v[8 * k8 + 6], v[8 * k8 + 7]);
            *reinterpret_cast<uint4*>(nxt + (size_t)((c0 >> 3) + k8) * kCsSlab) = w;
          }
        }
        tc_fence_before();
        fence_proxy_async_smem();
        mbar_arrive(&sh.img_ready[l + 1]);
      } else {
        // last layer: rows whose whole dependency cone was computed in this tile -> HBM
        const bool mine = inside && (i + h) >= L * h && (i + h) < 128 + 2 * h - L * h;
        for (int c0 = 0; c0 < ly.cout; c0 += 16) {
          float v[16];
          tmem_ld16(lane_addr + (uint32_t)c0, v);
          if (mine) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(ly.bias + c0) + q);
              float4 o = make_float4(v[4 * q] + b.x, v[4 * q + 1] + b.y, v[4 * q + 2] + b.z, v[4 * q + 3] + b.w);
              if (ly.act == FCL_ACT_RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
              else if (ly.act == FCL_ACT_TANH) { o.x = tanh_fast(o.x); o.y = tanh_fast(o.y); o.z = tanh_fast(o.z); o.w = tanh_fast(o.w); }
              if (p.residual) {
                const float4 r4 = __ldg(reinterpret_cast<const float4*>(p.residual + (size_t)g * p.ldr + c0) + q);
                o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;
              }
              reinterpret_cast<float4*>(p.out + (size_t)g * p.ldo + c0)[q] = o;
            }
          }
        }
      }
    }
  } else if (warp == 4) {
    // ------------------------------------------------ weight producer: layers in order, blocks [K stage][tap]
    if (elect_one()) {
      int it = 0;
      for (int l = 0; l < L; ++l) {
        const FclConvLayer& ly = p.layers[l];
        const uint32_t bb = (uint32_t)ly.cout * ly.kstage * 2u;
        const uint8_t* w = reinterpret_cast<const uint8_t*>(ly.w_packed);
        const int n = (ly.cin / ly.kstage) * taps;
        for (int b = 0; b < n; ++b, ++it) {
          const int s = it % kCsBStages;
          mbar_wait(&sh.b_empty[s], ((uint32_t)(it / kCsBStages) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&sh.b_full[s], bb);
          bulk_g2s(b_ring + (size_t)s * b_slot_bytes, w + (size_t)b * bb, bb, &sh.b_full[s]);
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------ MMA issuer
    if (elect_one()) {
      int it = 0;
      for (int l = 0; l < L; ++l) {
        const FclConvLayer& ly = p.layers[l];
        const uint32_t idesc = idesc_op_f32(128u, (uint32_t)ly.cout);
        const uint32_t b_lbo = (uint32_t)ly.cout * 16u;
        const int kchunks = ly.cin / ly.kstage, ksteps = ly.kstage / 16, slabs = ly.kstage / 8;
        mbar_wait(&sh.img_ready[l], 0);            // also orders this layer after the previous epilogue's TMEM reads
        tc_fence_after();
        const uint32_t a_img = smem_u32(img[l & 1]);
        bool first = true;
        for (int kc = 0; kc < kchunks; ++kc) {
          for (int t = 0; t < taps; ++t, ++it) {
            const int s = it % kCsBStages;
            mbar_wait(&sh.b_full[s], (uint32_t)(it / kCsBStages) & 1u);
            tc_fence_after();
            const uint32_t b_addr = smem_u32(b_ring + (size_t)s * b_slot_bytes);
            for (int k = 0; k < ksteps; ++k) {
              const uint64_t ad = smem_desc(a_img + (uint32_t)(kc * slabs + 2 * k) * kCsSlab + (uint32_t)t * 16u, kCsSlab, 128u);
              const uint64_t bd = smem_desc(b_addr + (uint32_t)k * 2u * b_lbo, b_lbo, 128u);
              mma_bf16_ss(tmem, ad, bd, idesc, first ? 0u : 1u);
              first = false;
            }
            mma_commit(&sh.b_empty[s]);
          }
        }
        mma_commit(&sh.accum);
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, tmem_cols);
}

// tiles of a ragged row space for the fused stack: stride S rows per tile, never crossing a segment
__global__ void __launch_bounds__(1024, 1)
conv_stack_tiles_kernel(FclConvStackTilesParams p) {
  __shared__ int warp_sums[32];
  __shared__ int carry;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < p.n_segs; base += 1024) {
    const int s = base + tid;
    int lo = 0, hi = 0;
    if (s < p.n_segs) { lo = p.seg_off[s]; hi = p.seg_off[s + 1]; }
    const int n = (hi - lo + p.stride - 1) / p.stride;
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
    const int first = carry + warp_sums[wid] + incl - n;
    for (int k = 0; k < n; ++k) {
      const int t = first + k;
      if (t < p.max_tiles) {
        p.tiles[4 * t] = lo + k * p.stride; p.tiles[4 * t + 1] = lo; p.tiles[4 * t + 2] = hi; p.tiles[4 * t + 3] = s;
      }
    }
    __syncthreads();
    if (tid == 1023) carry = first + n;
    __syncthreads();
  }
  if (tid == 0) *p.n_tiles = min(carry, p.max_tiles);
}

}  // namespace fcl

extern "C" int fcl_conv_stack_tiles(const FclConvStackTilesParams* p, void* stream) {
  using namespace fcl;
  FCL_REQUIRE(p && p->seg_off && p->tiles && p->n_tiles, "null pointer");
  FCL_REQUIRE(p->n_segs > 0 && p->max_tiles > 0 && p->stride > 0 && p->stride <= 128, "bad sizes");
  conv_stack_tiles_kernel<<<1, 1024, 0, as_stream(stream)>>>(*p);
  return check_launch("fcl_conv_stack_tiles");
}

extern "C" int fcl_conv_stack_bf16(const FclConvStackParams* p, void* stream) {
  using namespace fcl;
  FCL_REQUIRE(p && p->in && p->tiles && p->out, "null pointer");
  FCL_REQUIRE(p->n_layers >= 1 && p->n_layers <= FCL_MAX_STACK_LAYERS && p->n_tiles > 0, "bad sizes");
  FCL_REQUIRE((p->taps == 3 || p->taps == 5) && (128 - 2 * (p->taps / 2) * (p->n_layers - 1)) > 0, "taps must be 3 or 5");
  int max_c = 0, max_cout = 0;
  size_t b_slot = 0;
  for (int l = 0; l < p->n_layers; ++l) {
    const FclConvLayer& ly = p->layers[l];
    FCL_REQUIRE(ly.w_packed && ly.bias, "null layer pointer");
    FCL_REQUIRE(ly.kstage % 16 == 0 && ly.kstage <= 80 && ly.cin % ly.kstage == 0 && ly.cin % 16 == 0, "bad cin/kstage");
    FCL_REQUIRE(ly.cout % 16 == 0 && ly.cout <= 256, "cout must be a multiple of 16, <= 256");
    FCL_REQUIRE(l == 0 || ly.cin == p->layers[l - 1].cout, "layer sizes do not chain");
    max_c = ly.cin > max_c ? ly.cin : max_c;
    max_cout = ly.cout > max_cout ? ly.cout : max_cout;
    const size_t bb = (size_t)ly.cout * ly.kstage * 2;
    b_slot = bb > b_slot ? bb : b_slot;
  }
  FCL_REQUIRE(p->in_channels > 0 && p->in_channels <= p->layers[0].cin && p->in_channels % 16 == 0, "bad in_channels");
  FCL_REQUIRE(p->ld_in % 4 == 0 && p->ldo % 4 == 0 && (!p->residual || p->ldr % 4 == 0), "leading dims must be multiples of 4");
  const size_t img_bytes = (size_t)(max_c / 8) * kCsSlab;
  // The weight ring must hold enough bytes in flight to cover the L2 round trip (~2 k cycles x ~60 B/clk per SM):
  // as many stages as fit beside the two images (one CTA per SM when that is what it takes).
  if (img_bytes + 2 * b_slot > 216 * 1024) {
    set_error("fcl_conv_stack_bf16: %zu B shared memory needed (channels too wide to fuse)", img_bytes + 2 * b_slot);
    return FCL_EUNSUPPORTED;
  }
  // measured: the per-tile chain (load, then L x [MMA -> epilogue]) is latency-bound, so co-resident CTAs matter
  // more than a deep weight ring: 2 stages, as many CTAs per SM as shared memory / TMEM allow.
  int b_stages = p->b_stages > 0 ? p->b_stages : 2;
  b_stages = b_stages > kCsMaxBStages ? kCsMaxBStages : b_stages;
  const size_t smem = img_bytes + (size_t)b_stages * b_slot;
  if (smem > 216 * 1024) { set_error("fcl_conv_stack_bf16: %zu B shared memory needed (channels too wide to fuse)", smem); return FCL_EUNSUPPORTED; }
  if (int rc = ensure_dyn_smem(conv_stack_bf16_kernel, 216 * 1024, "fcl_conv_stack_bf16")) return rc;
  conv_stack_bf16_kernel<<<p->n_tiles, kCsThreads, smem, as_stream(stream)>>>(*p, (uint32_t)img_bytes, (uint32_t)b_slot,
                                                                               tmem_cols_pow2((uint32_t)max_cout), b_stages);
  return check_launch("fcl_conv_stack_bf16");
}
