/*
 * fcl_taco2.h -- C ABI of the B200 (sm_100a) FCL-taco2 inference path.
 *
 * The reference (Wendison/FCL-taco2) is pure Python on top of torch ops; it has
 * no FFI of its own. Each entry point below replaces the torch ops of one row
 * of SURVEY.md section 8(a); the reference lines it stands for are cited on each
 * declaration. INTEGRATION.md shows the ctypes stub a maintainer would add.
 *
 * Conventions
 *   - extern "C", POD structs, raw DEVICE pointers, explicit sizes; no torch types.
 *   - No allocation inside: the caller owns every buffer, including scratch.
 *   - Work is enqueued on `stream` (a cudaStream_t passed as void*); nothing
 *     synchronises the host.
 *   - Return 0 on success, a negative FCL_E* code otherwise; fcl_last_error()
 *     returns a thread-local message. Nothing throws.
 *   - "rows" are phonemes (ragged-packed over the utterances of a batch, an
 *     utterance = a contiguous row range) or mel frames (same, per utterance).
 *     Activations are row-major (rows, channels) fp32.
 *   - There is no CPU fallback: with no CUDA device every launch returns FCL_ECUDA.
 */
#ifndef FCL_TACO2_H_
#define FCL_TACO2_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FCL_ABI_VERSION 29

enum {
  FCL_OK = 0,
  FCL_EINVAL = -1,   /* bad argument (null pointer, unsupported size)            */
  FCL_ECUDA = -2,    /* CUDA runtime error (message has cudaGetErrorString)      */
  FCL_EUNSUPPORTED = -3
};

enum { FCL_ACT_NONE = 0, FCL_ACT_RELU = 1, FCL_ACT_TANH = 2 };

int fcl_abi_version(void);
/* Element format of the 16-bit GEMM operands (weights packed by the host, activation images written by the kernels) of
 * every tensor-core entry point (the `_bf16` suffix of their names is historical): 0 = IEEE fp16 (default build: 11-bit
 * significand, conversions saturate at +-65504), 1 = bfloat16 (library built with -DFCL_OPERANDS_BF16). The host packs
 * weights in the format this returns. */
int fcl_operand_format(void);
const char* fcl_last_error(void);
/* Number of SMs of the current device (grid sizing on the host side); <0 on error. */
int fcl_sm_count(void);
/* sizeof() of parameter struct number `which` (declaration order in this header, 0-based); a binding
 * in another language checks its own layout against it. -1 for an unknown index. */
int fcl_struct_size(int which);

/* ---------------------------------------------------------------- K0: length regulator
 * Replaces the Python loops of nets/teacher_training/e2e_tts_tacotron2_sa.py:665-671
 * (position table) and nets/modules/decoder_sa.py:619-630 (ragged gather order):
 * durations -> exclusive scan (frame offset of every phoneme row), per-utterance
 * frame offsets, the frame -> (row, step) index map, and a duration-descending
 * row order used to form decoder tiles. Integer work, bit-exact.
 */
typedef struct {
  int32_t n_rows;            /* P: phoneme rows in the batch                               */
  int32_t n_utts;            /* B                                                          */
  const int32_t* dur;        /* (P)  durations, >= 0 (0 rows produce no frames)            */
  const int32_t* utt_off;    /* (B+1) row offset of each utterance                         */
  int32_t* frame_off;        /* out (P+1) exclusive scan of dur                            */
  int32_t* utt_frame_off;    /* out (B+1) frame offset of each utterance                   */
  int32_t* order;            /* out (P)  rows sorted by duration, descending               */
  int32_t* totals;           /* out (2)  [0]=F total frames, [1]=max duration              */
  int32_t* ws;               /* optional scratch of fcl_len_reg_ws_ints(n_rows) int32: enables the multi-CTA path
                                (two launches: per-block sums + histogram, then scan + counting sort) for large batches;
                                NULL = one CTA                                              */
} FclLenRegParams;
#define FCL_MAX_DURATION 1023      /* largest supported duration (the host refuses larger ones; reference data cap is 50) */
int fcl_len_reg_scan(const FclLenRegParams* p, void* stream);
int fcl_len_reg_ws_ints(int32_t n_rows);

typedef struct {
  int32_t n_rows, n_utts, n_frames;
  const int32_t* frame_off;      /* (P+1) */
  const int32_t* utt_frame_off;  /* (B+1) */
  int32_t* frame_row;            /* out (F) row of each frame                                */
  int32_t* frame_step;           /* out (F) step m of each frame within its row              */
  int32_t* frame_seg_lo;         /* out (F) first frame of the frame's utterance             */
  int32_t* frame_seg_hi;         /* out (F) one past the last frame of the frame's utterance */
  float* position;               /* optional out (F): float(m)/float(d) (IEEE division)      */
} FclFrameMapParams;
int fcl_len_reg_frame_map(const FclFrameMapParams* p, void* stream);

/* ---------------------------------------------------------------- ragged conv1d / linear as GEMM
 * out[r, n] = act( sum_{t<taps} sum_{c<cin} A[r + t - taps/2, c] * W[t][c][n] + bias[n] ) (+ residual[r, n])
 * with A rows outside [seg_lo[r], seg_hi[r]) read as zero (zero halo per utterance,
 * never "pad and convolve": SURVEY.md 3.4). Replaces torch.nn.Conv1d (+ folded
 * eval-mode BatchNorm1d) + ReLU/Tanh of nets/modules/encoder_sa.py:135-140,
 * nets/modules/decoder_sa.py:274-286,632, the Conv1d of variance_predictor.py:86-87,
 * and torch.nn.Linear where taps == 1. With `gather` != NULL, A row r is
 * table row gather[r] (torch.nn.Embedding, encoder_sa.py:134).
 */
typedef struct {
  int32_t rows, cin, cout, taps;
  const float* a;            /* (rows, lda) or the embedding table when gather != NULL     */
  int32_t lda;
  const int64_t* gather;     /* optional (rows) int64 ids                                  */
  const int32_t* seg_lo;     /* (rows) or NULL when taps == 1                              */
  const int32_t* seg_hi;
  const float* w;            /* packed (taps, cin, cout) fp32                              */
  const float* bias;         /* optional (cout)                                            */
  const float* residual;     /* optional (rows, ldr)                                       */
  int32_t ldr;
  float* out;                /* (rows, ldo)                                                */
  int32_t ldo;
  int32_t act;               /* FCL_ACT_*                                                  */
} FclConvGemmParams;
int fcl_conv_gemm_f32(const FclConvGemmParams* p, void* stream);

/* Tensor-core form of the same contract (tcgen05, bf16 operands, fp32 accumulate in TMEM, fp32 activations
 * in HBM). `w_packed` holds bf16 weights pre-tiled as UMMA core matrices:
 *   [cout/ntile][cin/kstage][taps][kstage/8][ntile][8]   (see fcl_taco2_b200/pack.py: pack_conv_bf16)
 * so that one weight stage is one contiguous bulk copy. For taps > 1 the row tiles come from fcl_conv_tiles
 * (a tile never crosses an utterance; its input window is fetched once and the taps are descriptor shifts).
 */
typedef struct {
  int32_t n_segs;                 /* utterances                                              */
  int32_t max_tiles;              /* capacity of the maps: >= sum over segments of ceil(len/128) */
  int32_t halo;                   /* window halo rows on each side (>= taps/2 of every conv using the maps, <= 4) */
  const int32_t* seg_off;         /* (n_segs+1) row offsets of the segments                  */
  int32_t* seg_first_tile;        /* out (n_segs+1) scratch: first tile of each segment      */
  int32_t* tile_src;              /* out (max_tiles, 136): global row of each window row, -1 = zero */
  int32_t* tile_dst;              /* out (max_tiles, 128): global row of each output row, -1 = none */
  int32_t* n_tiles;               /* out (1): number of tiles                                */
} FclConvTilesParams;
int fcl_conv_tiles(const FclConvTilesParams* p, void* stream);

typedef struct {
  int32_t rows, cin, cout, taps;
  const float* a;
  int32_t lda;
  const int64_t* gather;     /* optional: A row r is table row gather[r] (embedding ids)   */
  const int32_t* row_gather; /* optional: A row r is row row_gather[r] of `a` (row permutation) */
  const int32_t* tile_src;   /* tile maps (fcl_conv_tiles); NULL => plain 128-row tiles (taps must be 1) */
  const int32_t* tile_dst;
  const int32_t* n_tiles_dev;/* optional device count of valid tiles (grid = n_tiles is an upper bound) */
  int32_t n_tiles;           /* tiles to launch when tile maps are given                   */
  int32_t map_halo;          /* halo the maps were built with                              */
  const void* w_packed;      /* bf16, layout above                                         */
  int32_t ntile;             /* output columns per CTA: multiple of 16, <= 256, divides cout */
  int32_t kstage;            /* K per pipeline stage: multiple of 16, <= 80, divides cin    */
  const float* bias;
  const float* residual;
  int32_t ldr;
  float* out;
  int32_t ldo;
  int32_t act;
  int32_t out_bf16;          /* != 0: `out` is bf16 (rows, ldo) instead of fp32             */
} FclConvGemmBf16Params;
int fcl_conv_gemm_bf16(const FclConvGemmBf16Params* p, void* stream);

/* Fused stack of k-tap conv layers (postnet: decoder_sa.py:274-286,632; encoder convs: encoder_sa.py:135-140):
 * all intermediate activations stay in shared memory as bf16 operand images; only the first input and the last
 * output touch HBM (halo recompute: a tile yields 128 - 2*(taps/2)*(n_layers-1) output rows).
 * Every layer: out = act(conv_k(in) + bias) with BatchNorm already folded; the last layer may add `residual`.
 * Usable when 2 * max(cin)/8 * 2176 B + 2 * max(cout*kstage*2) B fits in shared memory (FCL-taco2-S postnet/encoder).
 */
#define FCL_MAX_STACK_LAYERS 5
typedef struct {
  int32_t cin, cout, kstage, act;     /* kstage: multiple of 16 (<= 80) dividing cin; act: FCL_ACT_*          */
  const void* w_packed;               /* bf16 [cin/kstage][taps][kstage/8][cout][8] (pack_conv_bf16, ntile = cout) */
  const float* bias;                  /* (cout)                                                               */
} FclConvLayer;
typedef struct {
  int32_t n_segs, max_tiles, stride;  /* stride = output rows per tile                                        */
  const int32_t* seg_off;             /* (n_segs+1)                                                           */
  int32_t* tiles;                     /* out (max_tiles, 4): first output row, seg_lo, seg_hi, segment        */
  int32_t* n_tiles;                   /* out (1)                                                              */
} FclConvStackTilesParams;
int fcl_conv_stack_tiles(const FclConvStackTilesParams* p, void* stream);
typedef struct {
  int32_t n_layers, taps;
  FclConvLayer layers[FCL_MAX_STACK_LAYERS];
  const float* in;                    /* (rows, ld_in) fp32, or the embedding table when gather != NULL       */
  int32_t ld_in;
  int32_t in_channels;                /* real input width (<= layers[0].cin, which may be zero-padded to 64)  */
  int32_t b_stages;                   /* weight ring depth (0 = default 2)                                    */
  const int64_t* gather;
  const int32_t* tiles;               /* from fcl_conv_stack_tiles                                            */
  const int32_t* n_tiles_dev;         /* optional device tile count (grid = n_tiles is an upper bound)        */
  int32_t n_tiles;
  const float* residual;              /* optional (rows, ldr), added to the last layer                        */
  int32_t ldr;
  float* out;                         /* (rows, ldo)                                                          */
  int32_t ldo;
} FclConvStackParams;
int fcl_conv_stack_bf16(const FclConvStackParams* p, void* stream);

/* ---------------------------------------------------------------- image-to-image convolutions (tcgen05, cta_group::2, TMA)
 * Front-end form of the same convolutions (encoder_sa.py:134-140 conv stack, the LSTM input projection of
 * encoder_sa.py:143-146, variance_predictor.py:48-66,86-90 and the espnet DurationPredictor): activations travel
 * between layers as bf16 UMMA operand images in a PADDED row space, so a layer's input window is one tensor-map TMA
 * load per 64-channel stage (no fp32 round trip, no conversion pass) and tiles are dense:
 *   padded row of phoneme j of utterance u:  prow = utt_off[u] + gap*u + j      (gap zero rows between utterances:
 *   the zero halo of every k-tap conv; "zero halo per utterance", never pad-and-convolve -- SURVEY.md 3.4)
 *   image[c/8][prow + 4][c%8] bf16, (n_tiles*128 + 8) rows per 8-channel slab; rows that are not phonemes are ZERO
 *   (every producer writes them), so consumers need no per-utterance logic at all.
 * A CTA PAIR (2-CTA cluster) computes two 128-row tiles with one M=256 tcgen05.mma stream; each CTA loads its own
 * input window and HALF of every weight stage (tensor-map TMA with .cta_group::2 completion on the leader's mbarrier).
 * The grid is persistent (pairs walk super-tiles), accumulators live in TMEM (double-buffered when <= 256 columns).
 */
typedef struct {
  int32_t n_utts, n_rows, gap;   /* gap >= taps/2 of every conv that reads the images (2 for k <= 5)             */
  const int32_t* utt_off;        /* (B+1) row offsets (processing order)                                          */
  int32_t n_tiles;               /* ceil((n_rows + gap*(n_utts-1)) / 128), computed by the caller                 */
  int32_t* prow_src;             /* out (n_tiles*128): original row of every padded row, -1 = gap / tail          */
  int32_t* prow_off;             /* out (B+1): padded row of the first phoneme of every utterance                 */
} FclPadRowsParams;
int fcl_pad_rows(const FclPadRowsParams* p, void* stream);

typedef struct {
  int32_t n_tiles, chans;        /* chans % 8 == 0                                                                */
  const float* src;              /* (rows, ld) fp32, or the embedding table when gather != NULL                   */
  int32_t ld;
  const int64_t* gather;         /* optional: source row = gather[original row] (torch.nn.Embedding ids)          */
  const int32_t* prow_src;       /* (n_tiles*128) from fcl_pad_rows                                               */
  void* img;                     /* out bf16 [chans/8][n_tiles*128 + 8][8]                                        */
  int32_t src_chans;             /* channels present in `src` (multiple of 8); image channels beyond are zero. 0 = chans */
} FclRowsToImageParams;
int fcl_rows_to_image(const FclRowsToImageParams* p, void* stream);

enum {
  FCL_EPI_IMAGE = 0,        /* act(conv + bias) -> bf16 image                                                      */
  FCL_EPI_LN_IMAGE = 1,     /* LayerNorm_C(act(conv + bias)) * gamma + beta -> bf16 image (variance_predictor.py:52-64) */
  FCL_EPI_LN_HEAD = 2,      /* ... -> dot(., head_w) + head_b -> head_out[original row] (+ duration rounding)      */
  FCL_EPI_BLOCKED_F32 = 3,  /* conv + bias -> fp32, column-blocked by PADDED row: out[c/16][n_tiles*128][c%16]     */
  FCL_EPI_BLOCKED_F16 = 4,  /* same layout in fp16 (saturating): half the HBM bytes, 11-bit significand            */
  FCL_EPI_ROWS_F32 = 5      /* act(conv + bias) (+ residual) -> fp32 rows by ORIGINAL row: out_rows[row][c], c < out_chans
                               (last postnet layer + `before` residual: decoder_sa.py:285,632)                      */
};
typedef struct {
  int32_t n_tiles, cin, cout, taps;   /* cin % 64 == 0; taps in {1,3,5}                                           */
  int32_t nb;                         /* MMA N: multiple of 64, <= 256, divides cout (pack.py: pack_conv_pair)     */
  int32_t act, epi;                   /* FCL_ACT_*, FCL_EPI_*                                                      */
  const void* in_img;                 /* bf16 [cin/8][n_tiles*128 + 8][8]                                          */
  const void* w_packed;               /* bf16 [cout/nb][cin/64][taps][2 halves][8][nb/2][8]                        */
  const float* bias;                  /* optional (cout)                                                           */
  const int32_t* prow_src;            /* (n_tiles*128)                                                             */
  void* out_img;                      /* FCL_EPI_IMAGE / LN_IMAGE: bf16 [cout/8][n_tiles*128 + 8][8]               */
  float* out_blk;                     /* FCL_EPI_BLOCKED_F32 (float) / FCL_EPI_BLOCKED_F16 (__half)                */
  const float* gamma;                 /* LN epilogues: (cout), eps 1e-12 (espnet LayerNorm); cout <= 512           */
  const float* beta;
  const float* head_w;                /* FCL_EPI_LN_HEAD: (cout)                                                   */
  float head_b;
  float* head_out;                    /* (n_rows) by ORIGINAL row                                                  */
  int32_t* dur_out;                   /* optional: clamp(round_half_even(exp(head) - 1), 0, FCL_MAX_DURATION)      */
  int32_t n_pairs;                    /* CTA pairs to launch; 0 = min(SMs / 2, ceil(n_tiles / 2))                  */
  int64_t* trace;                     /* optional debug timeline of CTA 0: [0] = count (zero it), then (id, clock) pairs */
  int32_t trace_cap;                  /* capacity in records                                                        */
  float* out_rows;                    /* FCL_EPI_ROWS_F32: (n_rows, ldo) fp32                                       */
  int32_t ldo, out_chans;             /* out_chans <= cout, multiple of 4 (cout may be zero-padded to a multiple of 64) */
  const float* residual;              /* optional (n_rows, ldr), added after the activation                         */
  int32_t ldr;
} FclConvImgParams;
int fcl_conv_img_bf16(const FclConvImgParams* p, void* stream);

/* ---------------------------------------------------------------- LayerNorm (+ optional head)
 * y = LayerNorm_C(x) * gamma + beta, eps 1e-12 (espnet LayerNorm, variance_predictor.py:62).
 * If head_w != NULL: head[r] = dot(y[r], head_w) + head_b (Linear(C,1), variance_predictor.py:90)
 * and, if dur_out != NULL, dur_out[r] = clamp(round_half_even(exp(head[r]) - 1), 0, FCL_MAX_DURATION)
 * (espnet DurationPredictor.inference). `y` may be NULL when only the head is wanted.
 */
typedef struct {
  int32_t rows, chans;
  const float* x;            /* (rows, chans) */
  const float* gamma;
  const float* beta;
  float* y;                  /* optional (rows, chans) */
  const float* head_w;       /* optional (chans) */
  float head_b;
  float* head_out;           /* optional (rows) */
  int32_t* dur_out;          /* optional (rows) */
} FclLayerNormParams;
int fcl_layernorm_f32(const FclLayerNormParams* p, void* stream);

/* ---------------------------------------------------------------- pitch/energy embed + add
 * hn[r, e] = h[r, e] + sum_j wp[e][j]*pitch[r+j-4] + bp[e] + sum_j we[e][j]*energy[r+j-4] + be[e]
 * (Conv1d(1,E,k=9,pad=4) over the phoneme axis of each utterance,
 * e2e_tts_tacotron2_sa.py:435-443,657-658; add: decoder_sa.py:570-571).
 */
typedef struct {
  int32_t rows, chans, taps;
  const float* h;            /* (rows, chans) */
  const float* pitch;        /* (rows) */
  const float* energy;       /* (rows) */
  const int32_t* seg_lo;
  const int32_t* seg_hi;
  const float* wp;           /* (chans, taps) */
  const float* bp;
  const float* we;
  const float* be;
  float* hn;                 /* out (rows, chans); may be NULL when `img` is given */
  const int32_t* order;      /* optional (rows): with `img`, the decoder's duration-sorted row order (fcl_len_reg_scan) */
  void* img;                 /* optional out: hn as the decoder's 16-bit operand image [ceil(rows/128)][chans/8][128][8],
                                image row (tile*128 + i) = hn row order[tile*128 + i] (what fcl_pack_rows_bf16 would
                                produce from hn), written directly: no fp32 round trip through HBM */
} FclEmbedAddParams;
int fcl_embed_add_f32(const FclEmbedAddParams* p, void* stream);

/* ---------------------------------------------------------------- encoder BiLSTM recurrence
 * torch.nn.LSTM(E, E/2, 1, bidirectional) of nets/modules/encoder_sa.py:96-100,143-146
 * on every utterance of the batch independently. The input projection (x W_ih^T + b_ih + b_hh,
 * both directions) is done beforehand with fcl_conv_gemm_f32 into `gx`.
 */
typedef struct {
  int32_t n_utts, hidden;    /* hidden = E/2 per direction                                 */
  const int32_t* utt_off;    /* (B+1)                                                      */
  const float* gx;           /* (P, 2, hidden*4) gate-interleaved: col = dir*4H + unit*4 + gate(i,f,g,o) */
  const float* whh;          /* (2, hidden, hidden*4) : [dir][k][unit*4+gate] = W_hh[gate*H+unit][k]      */
  float* out;                /* (P, 2*hidden): [fwd | bwd]                                 */
  int32_t group;             /* utterances per CTA: 1 or 8                                 */
} FclBiLstmParams;
int fcl_bilstm_f32(const FclBiLstmParams* p, void* stream);

/* Tensor-core form (tcgen05): tiles of 128 utterances x direction, h kept on chip as a bf16 operand image,
 * W_hh streamed through a bulk-copy ring. Utterances should be ordered longest-first (the batch planner does).
 *   gx         : (P, 2, 4*hidden) bf16, gate-interleaved (fcl_conv_gemm_bf16 with out_bf16 = 1)
 *   whh_packed : bf16 [dir][4*hidden/256][hidden/64] blocks of [8][256][8] (pack.py: pack_bilstm_whh_bf16)
 *   c_ws       : scratch, ceil(n_utts/tile_utts) * 2 * hidden * 128 floats
 *   tile_utts  : utterances per CTA tile (32, 64 or 128): the recurrence is a latency/MUFU-bound serial chain,
 *                so smaller tiles spread it over more SMs (the MMA is M=128 either way)
 */
typedef struct {
  int32_t n_utts, hidden, tile_utts;
  const int32_t* utt_off;
  const void* gx;            /* bf16 rows (see above), or NULL when gx_blk is given                              */
  const void* whh_packed;
  float* c_ws;
  float* out;                /* (P, 2*hidden) fp32: [fwd | bwd] */
  const float* gx_blk;       /* optional: fp32 input projection, column-blocked by PADDED row
                                [8*hidden/16][gx_rows][16] (fcl_conv_img_bf16, FCL_EPI_BLOCKED_F32)               */
  const int32_t* prow_off;   /* (B+1) padded row of each utterance's first phoneme (with gx_blk)                  */
  int32_t gx_rows;           /* n_tiles*128 (with gx_blk)                                                         */
  int32_t gx_blk_half;       /* != 0: gx_blk holds fp16 (FCL_EPI_BLOCKED_F16) instead of fp32                    */
} FclBiLstmBf16Params;
int fcl_bilstm_bf16(const FclBiLstmBf16Params* p, void* stream);

/* ---------------------------------------------------------------- K4: persistent decoder
 * The step loop of nets/modules/decoder_sa.py:577-617 (Prenet :146-158 with its always-on
 * dropout, ZoneOutCell :63-96 around torch.nn.LSTMCell, feat_out :398) fused with the ragged
 * gather of :619-630: every CTA owns a tile of duration-sorted rows, keeps z/c state on chip
 * for all steps and stores frame (row, m) straight to before[frame_off[row] + m] for m < d.
 * The step-invariant terms are hoisted (gemm beforehand):
 *   g0h = hn W_ih0[:, :E]^T + b_ih0 + b_hh0   (P, 4H) gate-interleaved
 *   y0h = hn W_feat[:, H:]^T                  (P, O)
 */
typedef struct {
  int32_t n_rows, eunits, dunits, prenet_units, odim;
  const int32_t* order;      /* (P) duration-descending row order (fcl_len_reg_scan)      */
  const int32_t* dur;        /* (P)                                                        */
  const int32_t* frame_off;  /* (P+1)                                                      */
  const int32_t* row_utt;    /* (P) dropout key: utterance index of the row               */
  const int32_t* row_phone;  /* (P) dropout key: phoneme index within the utterance       */
  const float* g0h;          /* (P, 4H)                                                    */
  const float* y0h;          /* (P, O)                                                     */
  const float* wp0;          /* (O, U)   = prenet.0 weight^T                               */
  const float* bp0;          /* (U)                                                        */
  const float* wp1;          /* (U, U)   = prenet.1 weight^T                               */
  const float* bp1;
  const float* w0;           /* (U + H, 4H) rows [prenet part of W_ih0 ; W_hh0]^T, gate-interleaved cols */
  const float* wpos;         /* (4H)  position column of W_ih0, gate-interleaved           */
  const float* w1;           /* (2H, 4H) rows [W_ih1 ; W_hh1]^T, gate-interleaved cols     */
  const float* b1;           /* (4H)  b_ih1 + b_hh1, gate-interleaved                      */
  const float* wf;           /* (H, O) = W_feat[:, :H]^T                                   */
  float* cstate;             /* scratch (2, P, H) cell states                              */
  float* before;             /* out (F, O)                                                 */
  float zoneout;
  float dropout_p;           /* 0 => prenet dropout off                                    */
  uint64_t dropout_seed;
  int32_t tile_rows;         /* 16 or 32                                                   */
  const float* tf_y;         /* optional (F, O): TEACHER FORCING (decoder_sa.py:431-542, `prev_out = y`): the input of
                                step m of a row is its ground-truth frame m-1 instead of the frame it generated   */
} FclDecoderParams;
int fcl_decoder_f32(const FclDecoderParams* p, void* stream);

/* Row gather + bf16 pack into the UMMA operand image the tensor-core decoder streams:
 * dst[tile][cols/8][128][8] (bf16) with dst row (tile*128 + i) = src row order[tile*128 + i]; rows beyond
 * n_rows are zero. Used for the encoder state h (+ pitch/energy embeddings) in duration-sorted order.
 */
typedef struct {
  int32_t n_rows, cols;
  const float* src;          /* (n_rows, ld) fp32 */
  int32_t ld;
  const int32_t* order;      /* (n_rows) */
  void* dst;                 /* bf16, ceil(n_rows/128) * cols * 128 elements */
} FclPackRowsParams;
int fcl_pack_rows_bf16(const FclPackRowsParams* p, void* stream);

/* Tensor-core form of K4 (tcgen05; bf16 operands, fp32 accumulators in TMEM, fp32 cell state).
 * Tiles are 128 duration-sorted rows; a persistent grid of `n_slots` CTAs walks the tiles.
 *   hn_img        : encoder state of every tile as a bf16 operand image (fcl_pack_rows_bf16); it is a K-slice
 *                   of the cell-0 and feat_out GEMMs (nothing is hoisted into per-row fp32 tables)
 *   w_stream      : bf16 weights pre-tiled as UMMA core matrices in consumption order
 *                   prenet.0 (K padded 80->128) | prenet.1 | cell 0 [W_ih0 prenet part ; W_ih0 h part ; W_hh0] |
 *                   cell 1 [W_ih1 ; W_hh1] | feat_out [z part ; h part]  (fcl_taco2_b200/pack.py: pack_decoder_stream)
 *   act_priv / act_shared / c_ws : scratch, sizes from fcl_decoder_bf16_workspace()
 *   n_slots       : CTAs launched (<= SM count: the kernel is persistent, one CTA per SM); tiles are assigned to
 *                   the n_slots / group groups by fcl_decoder_schedule
 */
typedef struct {
  int32_t n_rows, n_tiles, n_slots, eunits, dunits, prenet_units, odim;
  const int32_t* order;
  const int32_t* dur;
  const int32_t* frame_off;
  const int32_t* row_utt;
  const int32_t* row_phone;
  const void* hn_img;
  const void* w_stream;
  const float* bp0;          /* (U) prenet.0 bias                                          */
  const float* bp1;          /* (U) prenet.1 bias                                          */
  const float* wpos;         /* (4H) gate-interleaved position column of W_ih0             */
  const float* b0;           /* (4H) gate-interleaved b_ih0 + b_hh0                        */
  const float* b1;           /* (4H) gate-interleaved b_ih1 + b_hh1                        */
  int32_t group;             /* CTAs that split the gate columns of ONE tile (1 = a tile per CTA). With few tiles
                                (small batch, single utterance) a group of g CTAs streams 1/g of the LSTM weights each
                                and exchanges z through L2 with two counter barriers per step.               */
  void* act_priv;            /* n_slots * priv_bytes_per_cta                               */
  void* act_shared;          /* (n_slots / group) * shared_bytes_per_group                 */
  float* c_ws;               /* n_slots * c_floats_per_cta                                 */
  int32_t* group_sync;       /* (n_slots / group) * 2 counters (zeroed by the launcher)    */
  float* before;             /* out (F, odim)                                              */
  float zoneout;
  float dropout_p;
  uint64_t dropout_seed;
  const int32_t* tile_slot;  /* (n_tiles) CTA slot that processes each tile (fcl_decoder_schedule)            */
  const int32_t* tile_rank;  /* (n_tiles) position of the tile in its slot's list                             */
  int64_t* trace;            /* optional debug timeline of CTA 0: [0] = count (zero it), then (id, clock) pairs */
  int32_t trace_cap;         /* capacity in records                                         */
  int32_t inflight;          /* fcl_decoder_bf16_pair: super-tiles a CTA pair keeps in flight (1 or 2; 0 = 1). With 2,
                                act_priv and c_ws hold TWO blocks per CTA (2 * priv_bytes_per_cta, 2 * c_floats_per_cta). */
  const void* tf_x1;         /* optional (F, prenet_units) 16-bit operand rows from fcl_prenet0_tf: TEACHER FORCING
                                (decoder_sa.py:431-542): row f holds dropout(relu(prenet.0(y_f))), the prenet.0 output
                                that feeds the step after frame f; the composed feat_out->prenet.0 chunk is skipped.
                                Supported by fcl_decoder_bf16 and fcl_decoder_bf16_pair_v1.                         */
} FclDecoderBf16Params;
/* prenet.0 of the GROUND-TRUTH frames for the teacher-forced decoder (decoder_sa.py:146-158 applied to `prev_out = y`):
 * x1[f] = dropout(relu(y[f] Wp0^T + b)), keyed like the inference path (utterance, phoneme, step + 1, layer 0), written
 * as 16-bit operand rows. HBM-bound row kernel on CUDA cores (80 x 256 MACs per frame). */
typedef struct {
  int32_t n_frames, odim, prenet_units;
  const float* y;            /* (F, odim) ground-truth frames, ragged-packed like the decoder's output            */
  const int32_t* frame_row;  /* (F) phoneme row of each frame   (fcl_len_reg_frame_map)                           */
  const int32_t* frame_step; /* (F) step of each frame within its row                                              */
  const int32_t* row_utt;    /* (P) dropout keys                                                                   */
  const int32_t* row_phone;
  const float* wp0;          /* (odim, U) = prenet.0 weight^T                                                      */
  const float* bp0;          /* (U)                                                                                */
  float dropout_p;
  uint64_t dropout_seed;
  void* x1;                  /* out (F, U) 16-bit operand format                                                   */
} FclPrenet0TfParams;
int fcl_prenet0_tf(const FclPrenet0TfParams* p, void* stream);

/* Longest-processing-time assignment of the duration-sorted tiles to the persistent CTAs (tile cost = its
 * step count + 1): every CTA ends at about the same time. One warp, ~20 us. */
typedef struct {
  int32_t n_rows, n_tiles, n_slots;   /* n_tiles = number of scheduled units = ceil(n_rows / unit_rows)  */
  int32_t unit_rows;         /* rows per scheduled unit: 128 (a tile) or 256 (a super-tile of a CTA pair) */
  const int32_t* order;      /* (P) duration-descending row order */
  const int32_t* dur;        /* (P) */
  int32_t* tile_slot;        /* out (n_tiles) */
  int32_t* tile_rank;        /* out (n_tiles) */
} FclDecoderScheduleParams;
int fcl_decoder_schedule(const FclDecoderScheduleParams* p, void* stream);
int fcl_decoder_bf16_workspace(int32_t prenet_units, int32_t dunits, int64_t* priv_bytes_per_cta,
                               int64_t* shared_bytes_per_group, int64_t* c_floats_per_cta);
int fcl_decoder_bf16(const FclDecoderBf16Params* p, void* stream);
/* cta_group::2 variant: n_slots CTAs = n_slots/2 pairs (2-CTA clusters); each pair walks SUPER-tiles (tiles 2j, 2j+1)
 * with one M=256 MMA stream and every weight stage split between the two SMs. Same parameters, except:
 * `group` is ignored, tile_slot/tile_rank are indexed by super-tile (fcl_decoder_schedule with unit_rows = 256,
 * n_slots = number of pairs), act_shared holds n_slots * shared_bytes_per_group, and w_stream uses the pair packing
 * (pack.py: pack_decoder_stream(pair=True): each stage block = [half 0][half 1], feat_out columns padded to 128). */
/* act_priv / c_ws are indexed with TWO blocks per CTA in this kernel whatever `inflight` is. */
int fcl_decoder_bf16_pair(const FclDecoderBf16Params* p, void* stream);
/* Round-1 form of the same kernel (one super-tile per pair at a time, per-role tile loops instead of the slot scheduler):
 * same parameters and results; `inflight` is ignored and act_priv / c_ws are indexed with ONE block per CTA.
 * The engine's default for large batches. When they fit (S-sized models: 14 KB) the epilogue constants (gate biases,
 * position column, prenet biases) are staged in shared memory behind the ring. */
int fcl_decoder_bf16_pair_v1(const FclDecoderBf16Params* p, void* stream);

/* ---------------------------------------------------------------- multi-GPU: peer-memory gather plumbing
 * The path shards by utterance with no collective on the hot path; the one exchange is the final
 * gather of the ragged mels to the root GPU. These entry points move it onto the copy engines: the root allocates
 * receive buffers (fcl_peer_alloc) and exports them (fcl_ipc_export, a 64-byte cudaIpcMemHandle_t); every other rank
 * maps them (fcl_ipc_open, from ITS device) and pushes finished mels with fcl_copy_async on a side stream -- DMA over
 * NVLink, no SM taken from the next pass. Hand-shakes are 32-bit flags in exported memory: fcl_write_flags stores
 * `value` into n flags, fcl_wait_flags makes `stream` wait until n flags equal `expect` (a one-warp polling kernel).
 * No reference counterpart (the reference decodes one utterance at a time on one device: tts.py:655-674).
 */
int fcl_peer_alloc(int64_t bytes, void** dptr);
int fcl_peer_free(void* dptr);
int fcl_ipc_export(void* dptr, uint8_t* handle64);
int fcl_ipc_open(const uint8_t* handle64, void** dptr);
int fcl_ipc_close(void* dptr);
int fcl_copy_async(void* dst, const void* src, int64_t bytes, void* stream);
int fcl_wait_flags(const int32_t* flags, int32_t n, int32_t expect, void* stream);
int fcl_write_flags(int32_t* flags, int32_t n, int32_t value, void* stream);

#ifdef __cplusplus
}
#endif
#endif  /* FCL_TACO2_H_ */
