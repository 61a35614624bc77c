"""Device-side orchestration of one batched `inference()` pass.

Python only sequences kernel launches through the C ABI (fcl_taco2_b200._lib);
torch supplies device memory and the stream. Every arithmetic step is a kernel
of libfcl_taco2.so -- there is no torch op on activations and no CPU fallback.

Pass order (reference: nets/teacher_training/e2e_tts_tacotron2_sa.py:624-683):
  encoder convs (embedding gather fused) -> BiLSTM -> duration/pitch/energy predictors
  -> pitch/energy embed add -> length regulator -> decoder (hoisted terms, persistent loop)
  -> frame map -> postnet (+ residual).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib
from ._lib import ACT_NONE, ACT_RELU, ACT_TANH, dptr
from .hparams import HParams
from .plan import output_chunks, BatchPlan


_ELEM_SIZE = {torch.float32: 4, torch.int32: 4, torch.bfloat16: 2, torch.float16: 2, torch.int64: 8, torch.uint8: 1}


@dataclass
class BatchResult:
    out: torch.Tensor                  # (F, odim) fp32, frames of all utterances in processing order
    utt_frame_off: np.ndarray          # (B+1) frame offsets in processing order (host)
    perm: np.ndarray                   # processing order -> caller index
    extras: dict = field(default_factory=dict)

    def per_utterance(self):
        """-> list of (L_i, odim) views in the caller's order."""
        outs = [None] * len(self.perm)
        for k, i in enumerate(self.perm):
            outs[int(i)] = self.out[int(self.utt_frame_off[k]): int(self.utt_frame_off[k + 1])]
        return outs


class Engine:
    def __init__(self, hp: HParams, packed: dict, device, precision: str = "fp32", bf16_gemms=None,
                 bf16_decoder: bool = True):
        """precision 'bf16': tcgen05 kernels. `bf16_gemms` (iterable of pack.GEMM_KEYS, default all) and
        `bf16_decoder` select which parts use them -- used by the error-budget diagnostics."""
        hp.validate()
        if precision not in ("fp32", "fp16", "bf16"):
            raise ValueError(f"unknown precision {precision!r}")
        # "fp16" = the 16-bit tensor-core path; "bf16" is its legacy alias. Which 16-bit operand format the kernels use
        # is a build property of the library (fp16 by default: csrc/umma.cuh), reported as `operand_format`.
        self.hp, self.precision = hp, ("fp32" if precision == "fp32" else "fp16")
        self.operand_format = _lib.operand_format() if precision != "fp32" else "fp32"
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.FclError("the B200 path runs on CUDA devices only (no CPU fallback)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        _lib.load()
        with torch.cuda.device(self.device):      # fcl_sm_count() and the kernels' per-device attributes use the CURRENT device
            self._init_device_state(hp, packed, precision, bf16_gemms, bf16_decoder)

    def _init_device_state(self, hp, packed, precision, bf16_gemms, bf16_decoder):
        self.w = {k: v.to(self.device) for k, v in packed.items()}
        self.wb = {}
        if precision != "fp32":
            from . import pack as _pack
            self.op_dtype = _pack.op_dtype()
            self.wb = {k: (t.to(self.device), nt, ks) for k, (t, nt, ks) in _pack.pack_bf16(packed).items()
                       if bf16_gemms is None or k in bf16_gemms}
            self.bf16_decoder = bf16_decoder
            # front end as image-to-image convolutions (csrc/conv_img_bf16.cu) when every GEMM runs in bf16
            self.wi = {k: (t.to(self.device), nb) for k, (t, nb) in _pack.pack_img(packed).items()} if bf16_gemms is None else {}
            # postnet layers for the same kernel (channels zero-padded to multiples of 64): (weights, nb, cin, cout, bias)
            self.wpost = {k: (t.to(self.device), nb, ci, co, b.to(self.device))
                          for k, (t, nb, ci, co, b) in _pack.pack_img_postnet(packed).items()} if bf16_gemms is None else {}
            self.dec_stream = _pack.pack_decoder_stream(packed, hp).to(self.device)
            self.dec_stream_pair = _pack.pack_decoder_stream(packed, hp, pair=True).to(self.device)
            self.blstm_whh_bf16 = None
            if "blstm_wih" in self.wb and (4 * (hp.eunits // 2)) % 256 == 0:
                self.blstm_whh_bf16 = _pack.pack_bilstm_whh_bf16(packed).to(self.device)
            self.n_slots = _lib.load().fcl_sm_count()
            priv_b, shared_b, c_f = _lib.decoder_bf16_workspace(hp.prenet_units, hp.dunits)
            # two blocks per CTA: the pair kernel keeps two super-tiles in flight (x1|x2 images and cell state per slot)
            self.dec_act_priv = torch.empty((2 * self.n_slots * priv_b,), dtype=torch.uint8, device=self.device)
            self.dec_act_shared = torch.empty((self.n_slots * shared_b,), dtype=torch.uint8, device=self.device)
            self.dec_c_ws = torch.empty((2 * self.n_slots * c_f,), dtype=torch.float32, device=self.device)
            self.dec_group_sync = torch.zeros((2 * self.n_slots,), dtype=torch.int32, device=self.device)
        self.head_b = {s: float(packed[f"{s}_head_b"][0]) for s in ("dur", "pitch", "energy")}
        self.launches = 0
        self._arena, self._arena_views, self._arena_seq, self._in_pass = [], [], 0, False
        self._stream_handle = None
        self.force_group = 0             # decoder group size override (tests); 0 = choose from the tile count
        # cta_group::2 decoder variant: "v1" = one super-tile per CTA pair at a time (per-role tile loops); "v2" = slot
        # scheduler with `pair_inflight` super-tiles in flight. Both are bit-identical and tested; v1 is the faster one
        # (S batch 1024: v1 1.78 ms, v2 2.01 / 1.99 ms with 1 / 2 in flight -- profiles/r02_decoder_inflight.md)
        self.pair_kernel = "v1"
        self.max_pairs = None            # tests: fewer CTA pairs than SMs / 2, so that every pair walks many super-tiles
        self.pair_inflight = 2           # v2 only
        self.use_pair = None             # cta_group::2 decoder (CTA pairs): None = when every SM has a tile anyway; True / False force it
        self.skip_zero_durations = False # extension: phonemes with d = 0 produce no frames (the reference's inference asserts)
        self.use_img_convs = bool(getattr(self, "wi", None))   # padded-row-space image convolutions for encoder + predictors
        self.use_img_postnet = bool(getattr(self, "wpost", None))   # ... and for the postnet (five launches, any channel width)
        self.use_encoder_stack = False   # measured: with 256 channels only a 2-stage weight ring fits beside the images
        self.stage_events = None      # when a list: (stage, start_event, stop_event) appended per stage (bench.py)

    # ------------------------------------------------------------------ workspace arena
    def _buf(self, shape, dtype):
        """Scratch tensor for the current pass. Inside `run_uploaded` buffers are recycled by call order, so a
        steady-state pass performs no allocator calls at all (a cudaMalloc/cudaFree hiccup of the caching allocator
        between launches shows up as GPU idle time); outside a pass (unit tests) it is a plain allocation."""
        if not self._in_pass:
            return torch.empty(shape, dtype=dtype, device=self.device)
        i = self._arena_seq
        self._arena_seq += 1
        if i < len(self._arena_views):
            hit = self._arena_views[i]
            if hit is not None and hit[0] == shape and hit[1] is dtype:
                return hit[2]                                   # same shape as in the previous pass: no tensor ops at all
        else:
            self._arena.append(None)
            self._arena_views.append(None)
        n = 1
        for d in (shape if isinstance(shape, (tuple, list)) else (shape,)):
            n *= int(d)
        nbytes = n * _ELEM_SIZE[dtype]
        cur = self._arena[i]
        if cur is None or cur.numel() < nbytes:
            cur = torch.empty((max(nbytes, 256) * 5 // 4 + 255) // 256 * 256, dtype=torch.uint8, device=self.device)
            self._arena[i] = cur
        view = cur[:nbytes].view(dtype).view(shape)
        self._arena_views[i] = (shape, dtype, view)
        return view

    # ------------------------------------------------------------------ launch helpers
    def _stream(self):
        h = self._stream_handle                       # cached for the duration of a pass (set by _run_uploaded)
        return h if h is not None else torch.cuda.current_stream(self.device).cuda_stream

    def _call(self, name, params):
        _lib.call(name, params, self._stream())
        self.launches += 1

    def conv_gemm(self, a, w, bias, rows, cin, cout, taps, act, seg=None, gather=None, residual=None, out=None,
                  lda=None, key=None, row_gather=None, out_bf16=False):
        if out is None:
            out = self._buf((rows, cout), torch.float32)
        if key is not None and key in self.wb:
            wp, ntile, kstage = self.wb[key]
            tm = seg[2] if (seg is not None and taps > 1) else None      # tile maps of the row space
            p = _lib.ConvGemmBf16Params(rows=rows, cin=cin, cout=cout, taps=taps, a=dptr(a), lda=lda or cin,
                                        gather=dptr(gather), row_gather=dptr(row_gather),
                                        tile_src=dptr(tm["src"]) if tm else None,
                                        tile_dst=dptr(tm["dst"]) if tm else None,
                                        n_tiles_dev=dptr(tm["count"]) if tm else None,
                                        n_tiles=tm["max_tiles"] if tm else 0, map_halo=tm["halo"] if tm else 0,
                                        w_packed=dptr(wp), ntile=ntile,
                                        kstage=kstage, bias=dptr(bias), residual=dptr(residual), ldr=cout,
                                        out=dptr(out), ldo=cout, act=act, out_bf16=1 if out_bf16 else 0)
            self._call("fcl_conv_gemm_bf16", p)
            return out
        p = _lib.ConvGemmParams(rows=rows, cin=cin, cout=cout, taps=taps, a=dptr(a), lda=lda or cin,
                                gather=dptr(gather), seg_lo=dptr(seg[0]) if seg else None,
                                seg_hi=dptr(seg[1]) if seg else None, w=dptr(w), bias=dptr(bias),
                                residual=dptr(residual), ldr=cout, out=dptr(out), ldo=cout, act=act)
        self._call("fcl_conv_gemm_f32", p)
        return out

    def conv_tiles(self, seg_off, n_segs, max_tiles, halo=2):
        """Tile maps of a ragged row space for the tensor-core convolutions (None on the fp32 path)."""
        if self.precision == "fp32":
            return None
        dev = self.device
        first = self._buf((n_segs + 1,), torch.int32)
        src = self._buf((max_tiles, 136), torch.int32)
        dst = self._buf((max_tiles, 128), torch.int32)
        count = self._buf((1,), torch.int32)
        self._call("fcl_conv_tiles", _lib.ConvTilesParams(n_segs=n_segs, max_tiles=max_tiles, halo=halo,
                                                          seg_off=dptr(seg_off), seg_first_tile=dptr(first),
                                                          tile_src=dptr(src), tile_dst=dptr(dst), n_tiles=dptr(count)))
        return {"src": src, "dst": dst, "count": count, "max_tiles": max_tiles, "halo": halo, "first": first}

    def layernorm(self, x, g, b, y=None, head_w=None, head_b=0.0, head_out=None, dur_out=None):
        rows, chans = x.shape
        p = _lib.LayerNormParams(rows=rows, chans=chans, x=dptr(x), gamma=dptr(g), beta=dptr(b), y=dptr(y),
                                 head_w=dptr(head_w), head_b=head_b, head_out=dptr(head_out), dur_out=dptr(dur_out))
        self._call("fcl_layernorm_f32", p)

    # ------------------------------------------------------------------ padded row space / image convolutions
    def pad_rows(self, utt_off, n_utts, n_rows):
        """Padded row space of the batch (include/fcl_taco2.h): -> dict(n_tiles, rows_alloc, prow_src, prow_off)."""
        gap = _lib.PAD_GAP
        n_tiles = (n_rows + gap * (n_utts - 1) + 127) // 128
        prow_src = self._buf((n_tiles * 128,), torch.int32)
        prow_off = self._buf((n_utts + 1,), torch.int32)
        self._call("fcl_pad_rows", _lib.PadRowsParams(n_utts=n_utts, n_rows=n_rows, gap=gap, utt_off=dptr(utt_off),
                                                      n_tiles=n_tiles, prow_src=dptr(prow_src), prow_off=dptr(prow_off)))
        return {"n_tiles": n_tiles, "rows_alloc": n_tiles * 128 + 8, "prow_src": prow_src, "prow_off": prow_off}

    def rows_to_image(self, src, ld, chans, pad, gather=None, src_chans=0):
        img = self._buf((chans // 8 * pad["rows_alloc"] * 8,), self.op_dtype)
        self._call("fcl_rows_to_image", _lib.RowsToImageParams(n_tiles=pad["n_tiles"], chans=chans, src=dptr(src), ld=ld,
                                                               gather=dptr(gather), prow_src=dptr(pad["prow_src"]),
                                                               img=dptr(img), src_chans=src_chans))
        return img

    def conv_img(self, key, in_img, pad, cin, cout, taps, act, epi, bias=None, gamma=None, beta=None, head_w=None,
                 head_b=0.0, head_out=None, dur_out=None, weights=None, out_rows=None, out_chans=0, residual=None):
        """One image-to-image conv launch. -> the output image (EPI_IMAGE / EPI_LN_IMAGE), the blocked buffer
        (EPI_BLOCKED_*) or None (EPI_LN_HEAD: results are in head_out / dur_out; EPI_ROWS_F32: in out_rows)."""
        wp, nb = weights if weights is not None else self.wi[key]
        out_img = out_blk = None
        if epi in (_lib.EPI_IMAGE, _lib.EPI_LN_IMAGE):
            out_img = self._buf((cout // 8 * pad["rows_alloc"] * 8,), self.op_dtype)
        elif epi == _lib.EPI_BLOCKED_F32:
            out_blk = self._buf((cout // 16 * pad["n_tiles"] * 128 * 16,), torch.float32)
        elif epi == _lib.EPI_BLOCKED_F16:
            out_blk = self._buf((cout // 16 * pad["n_tiles"] * 128 * 16,), torch.float16)
        self._call("fcl_conv_img_bf16", _lib.ConvImgParams(
            n_tiles=pad["n_tiles"], cin=cin, cout=cout, taps=taps, nb=nb, act=act, epi=epi, in_img=dptr(in_img),
            w_packed=dptr(wp), bias=dptr(bias), prow_src=dptr(pad["prow_src"]), out_img=dptr(out_img), out_blk=dptr(out_blk),
            gamma=dptr(gamma), beta=dptr(beta), head_w=dptr(head_w), head_b=head_b, head_out=dptr(head_out),
            dur_out=dptr(dur_out), n_pairs=0, out_rows=dptr(out_rows), ldo=out_rows.shape[1] if out_rows is not None else 0,
            out_chans=out_chans, residual=dptr(residual), ldr=residual.shape[1] if residual is not None else 0))
        return out_img if out_img is not None else out_blk

    # ------------------------------------------------------------------ stages
    def encoder(self, ids, utt_off, seg, n_utts, lens=None):
        hp, w = self.hp, self.w
        P, E = ids.shape[0], hp.eunits
        x = None
        pad = seg[3] if len(seg) > 3 else None
        if pad is not None:
            # embedding -> image, three k5 convs image -> image, input projection -> blocked fp32, recurrence
            x = self.rows_to_image(w["embed"], hp.embed_dim, hp.embed_dim, pad, gather=ids)
            cin = hp.embed_dim
            for l in range(3):
                x = self.conv_img(f"enc_conv{l}", x, pad, cin, hp.econv_chans, 5, ACT_RELU, _lib.EPI_IMAGE,
                                  bias=w[f"enc_conv{l}_b"])
                cin = hp.econv_chans
            # gate pre-activations in fp16 (11-bit significand, |x| << 65504): half the HBM bytes of fp32, and the
            # write-back of this launch is what bounds it; bf16 here was the largest single rounding of the encoder
            gx = self.conv_img("blstm_wih", x, pad, hp.econv_chans, 4 * E, 1, ACT_NONE, _lib.EPI_BLOCKED_F16, bias=w["blstm_b"])
            return self._bilstm_bf16(None, utt_off, n_utts, P, gx_blk=gx, pad=pad)
        if self.precision != "fp32" and self.use_encoder_stack and all(f"enc_conv{l}" in self.wb for l in range(3)):
            if lens is None:
                lens = np.diff(utt_off.cpu().numpy().astype(np.int64))
            x = self.conv_stack([f"enc_conv{l}" for l in range(3)], [ACT_RELU] * 3, w["embed"], hp.embed_dim, P, utt_off,
                                n_utts, lambda s: int(((lens + s - 1) // s).sum()), gather=ids)
        if x is None:
            x = self._encoder_convs(ids, seg, P)
        return self._encoder_lstm(x, utt_off, n_utts, P)

    def _encoder_convs(self, ids, seg, P):
        hp, w = self.hp, self.w
        x = self.conv_gemm(w["embed"], w["enc_conv0_w"], w["enc_conv0_b"], P, hp.embed_dim, hp.econv_chans, 5,
                           ACT_RELU, seg=seg, gather=ids, key="enc_conv0")
        x = self.conv_gemm(x, w["enc_conv1_w"], w["enc_conv1_b"], P, hp.econv_chans, hp.econv_chans, 5, ACT_RELU, seg=seg,
                           key="enc_conv1")
        return self.conv_gemm(x, w["enc_conv2_w"], w["enc_conv2_b"], P, hp.econv_chans, hp.econv_chans, 5, ACT_RELU,
                              seg=seg, key="enc_conv2")

    def _encoder_lstm(self, x, utt_off, n_utts, P):
        hp, w = self.hp, self.w
        E = hp.eunits
        h = self._buf((P, E), torch.float32)
        if self.precision != "fp32" and getattr(self, "blstm_whh_bf16", None) is not None:
            gx = self._buf((P, 4 * E), self.op_dtype)
            self.conv_gemm(x, None, w["blstm_b"], P, hp.econv_chans, 4 * E, 1, ACT_NONE, key="blstm_wih", out=gx,
                           out_bf16=True)
            return self._bilstm_bf16(gx, utt_off, n_utts, P, h=h)
        gx = self.conv_gemm(x, w["blstm_wih"], w["blstm_b"], P, hp.econv_chans, 4 * E, 1, ACT_NONE, key="blstm_wih")
        p = _lib.BiLstmParams(n_utts=n_utts, hidden=E // 2, utt_off=dptr(utt_off), gx=dptr(gx), whh=dptr(w["blstm_whh"]),
                              out=dptr(h), group=8 if n_utts >= 8 else 1)
        self._call("fcl_bilstm_f32", p)
        return h

    def _bilstm_bf16(self, gx, utt_off, n_utts, P, h=None, gx_blk=None, pad=None):
        E = self.hp.eunits
        if h is None:
            h = self._buf((P, E), torch.float32)
        tile_utts = 32 if n_utts <= 32 * self.n_slots // 2 else 64 if n_utts <= 64 * self.n_slots // 2 else 128
        n_tiles = (n_utts + tile_utts - 1) // tile_utts
        c_ws = self._buf((n_tiles * 2 * (E // 2) * 128,), torch.float32)
        self._call("fcl_bilstm_bf16", _lib.BiLstmBf16Params(
            n_utts=n_utts, hidden=E // 2, tile_utts=tile_utts, utt_off=dptr(utt_off), gx=dptr(gx),
            whh_packed=dptr(self.blstm_whh_bf16), c_ws=dptr(c_ws), out=dptr(h), gx_blk=dptr(gx_blk),
            prow_off=dptr(pad["prow_off"]) if pad else None, gx_rows=pad["n_tiles"] * 128 if pad else 0,
            gx_blk_half=1 if (gx_blk is not None and gx_blk.dtype == torch.float16) else 0))
        return h

    def predictor(self, name, h, seg, want_dur=False):
        """variance_predictor.py:86-93 / espnet DurationPredictor. -> (head (P,), dur int32 (P,) or None)"""
        hp, w = self.hp, self.w
        P, C = h.shape[0], hp.predictor_chans
        pad = seg[3] if len(seg) > 3 else None
        if pad is not None:
            if "h_img" not in pad:                   # shared by the duration / pitch / energy predictors of a pass
                pad["h_img"] = self.rows_to_image(h, hp.eunits, hp.eunits, pad)
            x = self.conv_img(f"{name}_conv0", pad["h_img"], pad, hp.eunits, C, 3, ACT_RELU, _lib.EPI_LN_IMAGE,
                              bias=w[f"{name}_conv0_b"], gamma=w[f"{name}_ln0_g"], beta=w[f"{name}_ln0_b"])
            head = self._buf((P,), torch.float32)
            dur = self._buf((P,), torch.int32) if want_dur else None
            self.conv_img(f"{name}_conv1", x, pad, C, C, 3, ACT_RELU, _lib.EPI_LN_HEAD, bias=w[f"{name}_conv1_b"],
                          gamma=w[f"{name}_ln1_g"], beta=w[f"{name}_ln1_b"], head_w=w[f"{name}_head_w"],
                          head_b=self.head_b[name], head_out=head, dur_out=dur)
            return head, dur
        x = self.conv_gemm(h, w[f"{name}_conv0_w"], w[f"{name}_conv0_b"], P, hp.eunits, C, 3, ACT_RELU, seg=seg,
                           key=f"{name}_conv0")
        self.layernorm(x, w[f"{name}_ln0_g"], w[f"{name}_ln0_b"], y=x)
        x = self.conv_gemm(x, w[f"{name}_conv1_w"], w[f"{name}_conv1_b"], P, C, C, 3, ACT_RELU, seg=seg, key=f"{name}_conv1")
        head = self._buf((P,), torch.float32)
        dur = self._buf((P,), torch.int32) if want_dur else None
        self.layernorm(x, w[f"{name}_ln1_g"], w[f"{name}_ln1_b"], head_w=w[f"{name}_head_w"],
                       head_b=self.head_b[name], head_out=head, dur_out=dur)
        return head, dur

    def embed_add(self, h, pitch, energy, seg, order=None, want_rows=True):
        """hn = h + pitch_embed + energy_embed. With `order` (tensor-core decoder) the result is also / only written as
        the decoder's operand image in duration-sorted tile order. -> (hn rows or None, image or None)"""
        hp, w = self.hp, self.w
        P, E = h.shape[0], hp.eunits
        hn = self._buf(tuple(h.shape), torch.float32) if (want_rows or order is None) else None
        img = self._buf((((P + 127) // 128) * 128 * E,), self.op_dtype) if order is not None else None
        p = _lib.EmbedAddParams(rows=P, chans=E, taps=hp.embed_kernel, h=dptr(h), pitch=dptr(pitch),
                                energy=dptr(energy), seg_lo=dptr(seg[0]), seg_hi=dptr(seg[1]), wp=dptr(w["pemb_w"]),
                                bp=dptr(w["pemb_b"]), we=dptr(w["eemb_w"]), be=dptr(w["eemb_b"]), hn=dptr(hn),
                                order=dptr(order), img=dptr(img))
        self._call("fcl_embed_add_f32", p)
        return hn, img

    def len_reg_scan(self, dur, utt_off, n_utts):
        P = dur.shape[0]
        dev = self.device
        frame_off = self._buf((P + 1,), torch.int32)
        utt_frame_off = self._buf((n_utts + 1,), torch.int32)
        order = self._buf((P,), torch.int32)
        totals = self._buf((2,), torch.int32)
        ws = self._buf((_lib.load().fcl_len_reg_ws_ints(P),), torch.int32)       # multi-CTA scan + counting sort when P is large
        p = _lib.LenRegParams(n_rows=P, n_utts=n_utts, dur=dptr(dur), utt_off=dptr(utt_off), frame_off=dptr(frame_off),
                              utt_frame_off=dptr(utt_frame_off), order=dptr(order), totals=dptr(totals), ws=dptr(ws))
        self._call("fcl_len_reg_scan", p)
        return frame_off, utt_frame_off, order, totals

    def frame_map(self, frame_off, utt_frame_off, n_rows, n_utts, n_frames, want_position=False):
        dev = self.device
        buf = self._buf((4, max(n_frames, 1)), torch.int32)
        pos = self._buf((max(n_frames, 1),), torch.float32) if want_position else None
        p = _lib.FrameMapParams(n_rows=n_rows, n_utts=n_utts, n_frames=n_frames, frame_off=dptr(frame_off),
                                utt_frame_off=dptr(utt_frame_off), frame_row=dptr(buf[0]), frame_step=dptr(buf[1]),
                                frame_seg_lo=dptr(buf[2]), frame_seg_hi=dptr(buf[3]), position=dptr(pos))
        self._call("fcl_len_reg_frame_map", p)
        return buf, pos

    def decoder(self, hn, dur, frame_off, order, row_utt, row_phone, n_frames, zoneout, dropout_p, dropout_seed,
                tile_rows=None, schedule=None, tf=None, hn_img=None):
        """`tf` = (ys (F, O) fp32 ground-truth frames in output order, frame_row, frame_step): teacher forcing.
        `hn_img`: hn already packed as the tensor-core decoder's operand image (Engine.embed_add with `order`)."""
        hp, w = self.hp, self.w
        P, E, H, O = order.shape[0], hp.eunits, hp.dunits, hp.odim
        if self.precision != "fp32" and self.bf16_decoder:
            tf_x1 = None
            if tf is not None:
                # prenet.0 of the ground-truth frames for all steps at once (it does not depend on the recurrence)
                tf_x1 = self._buf((max(n_frames, 1), hp.prenet_units), self.op_dtype)
                self._call("fcl_prenet0_tf", _lib.Prenet0TfParams(
                    n_frames=n_frames, odim=O, prenet_units=hp.prenet_units, y=dptr(tf[0]), frame_row=dptr(tf[1]),
                    frame_step=dptr(tf[2]), row_utt=dptr(row_utt), row_phone=dptr(row_phone), wp0=dptr(w["dec_wp0"]),
                    bp0=dptr(w["dec_bp0"]), dropout_p=dropout_p, dropout_seed=dropout_seed, x1=dptr(tf_x1)))
            return self.decoder_bf16(hn, dur, frame_off, order, row_utt, row_phone, n_frames, zoneout, dropout_p,
                                     dropout_seed, schedule, tf_x1=tf_x1, hn_img=hn_img)
        with self.stage("decoder_hoist"):
            g0h = self.conv_gemm(hn, w["dec_g0h_w"], w["dec_g0h_b"], P, E, 4 * H, 1, ACT_NONE)
            y0h = self.conv_gemm(hn, w["dec_y0h_w"], None, P, E, O, 1, ACT_NONE)
        cstate = self._buf((2, P, H), torch.float32)
        before = self._buf((max(n_frames, 1), O), torch.float32)
        if tile_rows is None:
            tile_rows = 32 if H <= 512 else 16
        p = _lib.DecoderParams(n_rows=P, eunits=E, dunits=H, prenet_units=hp.prenet_units, odim=O, order=dptr(order),
                               dur=dptr(dur), frame_off=dptr(frame_off), row_utt=dptr(row_utt), row_phone=dptr(row_phone),
                               g0h=dptr(g0h), y0h=dptr(y0h), wp0=dptr(w["dec_wp0"]), bp0=dptr(w["dec_bp0"]),
                               wp1=dptr(w["dec_wp1"]), bp1=dptr(w["dec_bp1"]), w0=dptr(w["dec_w0"]),
                               wpos=dptr(w["dec_wpos"]), w1=dptr(w["dec_w1"]), b1=dptr(w["dec_b1"]), wf=dptr(w["dec_wf"]),
                               cstate=dptr(cstate), before=dptr(before), zoneout=zoneout, dropout_p=dropout_p,
                               dropout_seed=dropout_seed, tile_rows=tile_rows, tf_y=dptr(tf[0]) if tf is not None else None)
        with self.stage("decoder_loop"):
            self._call("fcl_decoder_f32", p)
        return before

    def decoder_schedule(self, order, dur, P):
        """Group size + longest-processing-time tile assignment of the tensor-core decoder (integer work that only
        depends on the durations: runs on the side stream, off the critical path)."""
        H = self.hp.dunits
        n_tiles = (P + 127) // 128
        # few tiles (small batch / one utterance): groups of CTAs split each tile's gate columns so that every SM
        # streams only its share of the LSTM weights
        gate_chunks = 4 * H // 256
        group = 1
        for g in range(2, min(gate_chunks, 16, self.n_slots // max(n_tiles, 1)) + 1):
            if -(-gate_chunks // g) < -(-gate_chunks // group):      # fewer chunks per CTA
                group = g
        group = self.force_group or group
        use_pair = self.use_pair if self.use_pair is not None else n_tiles >= self.n_slots
        if use_pair and group == 1 and n_tiles >= 2:
            # cta_group::2: pairs of CTAs walk super-tiles of 256 rows
            n_super = (n_tiles + 1) // 2
            n_pairs = min(self.max_pairs or self.n_slots // 2, self.n_slots // 2, n_super)
            sched = self._buf((2, n_super), torch.int32)
            self._call("fcl_decoder_schedule", _lib.DecoderScheduleParams(n_rows=P, n_tiles=n_super, n_slots=n_pairs,
                                                                          unit_rows=256, order=dptr(order), dur=dptr(dur),
                                                                          tile_slot=dptr(sched[0]), tile_rank=dptr(sched[1])))
            return -1, n_pairs, 2 * n_pairs, sched          # group -1 marks pair mode
        n_groups = min(self.n_slots // group, n_tiles)
        n_slots = n_groups * group
        sched = self._buf((2, n_tiles), torch.int32)
        self._call("fcl_decoder_schedule", _lib.DecoderScheduleParams(n_rows=P, n_tiles=n_tiles, n_slots=n_groups,
                                                                      unit_rows=128, order=dptr(order), dur=dptr(dur),
                                                                      tile_slot=dptr(sched[0]), tile_rank=dptr(sched[1])))
        return group, n_groups, n_slots, sched

    def decoder_bf16(self, hn, dur, frame_off, order, row_utt, row_phone, n_frames, zoneout, dropout_p, dropout_seed,
                     schedule=None, tf_x1=None, hn_img=None):
        """Tensor-core decoder: h packed as a bf16 operand image in duration-sorted tile order, then the
        persistent tcgen05 loop."""
        hp, w = self.hp, self.w
        P, E, H, O = order.shape[0], hp.eunits, hp.dunits, hp.odim
        n_tiles = (P + 127) // 128
        if hn_img is None:
            with self.stage("decoder_hoist"):
                hn_img = self._buf((n_tiles * 128 * E,), self.op_dtype)
                self._call("fcl_pack_rows_bf16", _lib.PackRowsParams(n_rows=P, cols=E, src=dptr(hn), ld=E, order=dptr(order),
                                                                     dst=dptr(hn_img)))
        group, n_groups, n_slots, sched = schedule if schedule is not None else self.decoder_schedule(order, dur, P)
        before = self._buf((max(n_frames, 1), O), torch.float32)
        trace = getattr(self, "dec_trace", None)
        p = _lib.DecoderBf16Params(n_rows=P, n_tiles=n_tiles, n_slots=n_slots, eunits=E, dunits=H,
                                   prenet_units=hp.prenet_units, odim=O, order=dptr(order), dur=dptr(dur),
                                   frame_off=dptr(frame_off), row_utt=dptr(row_utt), row_phone=dptr(row_phone),
                                   hn_img=dptr(hn_img), w_stream=dptr(self.dec_stream_pair if group < 0 else self.dec_stream),
                                   bp0=dptr(w["dec_bp0"]), bp1=dptr(w["dec_bp1"]), wpos=dptr(w["dec_wpos"]),
                                   b0=dptr(w["dec_g0h_b"]), b1=dptr(w["dec_b1"]), group=max(group, 1),
                                   act_priv=dptr(self.dec_act_priv), act_shared=dptr(self.dec_act_shared),
                                   c_ws=dptr(self.dec_c_ws), group_sync=dptr(self.dec_group_sync), before=dptr(before),
                                   zoneout=zoneout,
                                   dropout_p=dropout_p, dropout_seed=dropout_seed, tile_slot=dptr(sched[0]),
                                   tile_rank=dptr(sched[1]), trace=dptr(trace),
                                   trace_cap=(-1 if getattr(self, "dec_prof", False) else (trace.numel() - 2) // 2) if trace is not None else 0,
                                   inflight=self.pair_inflight, tf_x1=dptr(tf_x1))
        with self.stage("decoder_loop"):
            self._call(self._pair_entry(tf_x1 is not None) if group < 0 else "fcl_decoder_bf16", p)
        return before

    def _pair_entry(self, teacher_forcing):
        if self.pair_kernel == "v2" and not teacher_forcing:
            return "fcl_decoder_bf16_pair"
        return "fcl_decoder_bf16_pair_v1"

    def conv_stack(self, keys, acts, x, ld_in, rows, seg_off, n_segs, max_len_sum_tiles, taps=5, gather=None,
                   residual=None, wkeys=None, final=False, out=None):
        """Fused conv stack (fcl_conv_stack_bf16) over the layers `keys`; returns None when the stack does not
        fit on chip (caller falls back to layer-by-layer fcl_conv_gemm_bf16)."""
        L = len(keys)
        stride = 128 - 2 * (taps // 2) * (L - 1)
        layers = (_lib.ConvLayer * _lib.MAX_STACK_LAYERS)()
        max_c, b_slot = 0, 0
        for l, key in enumerate(keys):
            wp, ntile, kstage = self.wb[(wkeys or keys)[l]]
            taps_l, cin, cout = self.w[key + "_w"].shape
            if ntile != cout or taps_l != taps:
                return None
            if l == 0:
                in_channels = cin
                cin = (cin + kstage - 1) // kstage * kstage        # zero-padded by the packer
            layers[l] = _lib.ConvLayer(cin=cin, cout=cout, kstage=kstage, act=acts[l], w_packed=dptr(wp),
                                       bias=dptr(self.w[key + "_b"]))
            max_c, b_slot = max(max_c, cin), max(b_slot, cout * kstage * 2)
        if (max_c // 8) * 2176 + 2 * b_slot > 216 * 1024:
            return None
        dev = self.device
        max_tiles = max_len_sum_tiles(stride)
        tiles = self._buf((max_tiles, 4), torch.int32)
        count = self._buf((1,), torch.int32)
        self._call("fcl_conv_stack_tiles", _lib.ConvStackTilesParams(n_segs=n_segs, max_tiles=max_tiles, stride=stride,
                                                                     seg_off=dptr(seg_off), tiles=dptr(tiles),
                                                                     n_tiles=dptr(count)))
        cout_last = layers[L - 1].cout
        if out is None:
            out = torch.empty((rows, cout_last), dtype=torch.float32, device=dev) if final else \
                self._buf((rows, cout_last), torch.float32)
        self._call("fcl_conv_stack_bf16", _lib.ConvStackParams(n_layers=L, taps=taps, layers=layers, in_=dptr(x), ld_in=ld_in,
                                                               in_channels=in_channels, b_stages=0, gather=dptr(gather), tiles=dptr(tiles), n_tiles_dev=dptr(count),
                                                               n_tiles=max_tiles, residual=dptr(residual), ldr=cout_last,
                                                               out=dptr(out), ldo=cout_last))
        return out

    def postnet(self, before, fseg, n_frames, chunks=None, chunk_cb=None):
        """`chunks` = [(first utt, last utt + 1, first frame, last frame + 1), ...] (processing order): the fused stack
        is launched per chunk and `chunk_cb(k, out, frame_lo, frame_hi)` is called right after chunk k is enqueued, so a
        consumer (the multi-GPU gather, a D2H copy) can start on finished frames while later chunks still compute."""
        hp, w = self.hp, self.w
        O, C = hp.odim, hp.postnet_chans
        if self.precision != "fp32" and self.use_img_postnet and len(fseg) > 3 and n_frames > 0:
            # five image-to-image launches in the padded FRAME space (decoder_sa.py:274-286): the mel input is converted
            # once (80 -> 128 zero-padded channels), activations stay bf16 images, the last layer writes fp32 rows and
            # adds the residual (decoder_sa.py:632)
            utt_frame_off, n_utts = fseg[3]
            pad = self.pad_rows(utt_frame_off, n_utts, n_frames)
            x = self.rows_to_image(before, O, self.wpost["post_conv0"][2], pad, src_chans=O)
            for l in range(4):
                wp, nb, ci, co, b = self.wpost[f"post_conv{l}"]
                x = self.conv_img(None, x, pad, ci, co, 5, ACT_TANH, _lib.EPI_IMAGE, bias=b, weights=(wp, nb))
            wp, nb, ci, co, b = self.wpost["post_conv4"]
            final = torch.empty((n_frames, O), dtype=torch.float32, device=self.device)     # returned to the caller
            self.conv_img(None, x, pad, ci, co, 5, ACT_NONE, _lib.EPI_ROWS_F32, bias=b, weights=(wp, nb), out_rows=final,
                          out_chans=O, residual=before)
            if chunks and chunk_cb is not None:
                for k, (u0, u1, f0, f1) in enumerate(chunks):          # one launch sequence: hand everything over at the end
                    chunk_cb(k, final, f0, f1)
            return final
        stack_ok = self.precision != "fp32" and len(fseg) > 3 and all(f"post_conv{l}" in self.wb for l in range(5))
        if stack_ok and chunks and chunk_cb is not None:
            utt_frame_off, n_utts = fseg[3]
            final = torch.empty((n_frames, O), dtype=torch.float32, device=self.device)
            ok = True
            for k, (u0, u1, f0, f1) in enumerate(chunks):
                r = self.conv_stack([f"post_conv{l}" for l in range(5)], [ACT_TANH] * 4 + [ACT_NONE], before, O, n_frames,
                                    utt_frame_off[u0:], u1 - u0, lambda s, f0=f0, f1=f1, n=u1 - u0: (f1 - f0 + s - 1) // s + n,
                                    residual=before, out=final,
                                    wkeys=[f"post_stack{l}" if f"post_stack{l}" in self.wb else f"post_conv{l}"
                                           for l in range(5)])
                if r is None:
                    ok = False
                    break
                chunk_cb(k, final, f0, f1)
            if ok:
                return final
        if stack_ok:
            utt_frame_off, n_utts = fseg[3]
            out = self.conv_stack([f"post_conv{l}" for l in range(5)], [ACT_TANH] * 4 + [ACT_NONE], before, O, n_frames,
                                  utt_frame_off, n_utts, lambda s: (n_frames + s - 1) // s + n_utts, residual=before,
                                  final=True,
                                  wkeys=[f"post_stack{l}" if f"post_stack{l}" in self.wb else f"post_conv{l}"
                                         for l in range(5)])
            if out is not None:
                if chunks and chunk_cb is not None:
                    for k, (u0, u1, f0, f1) in enumerate(chunks):
                        chunk_cb(k, out, f0, f1)
                return out
        x = self.conv_gemm(before, w["post_conv0_w"], w["post_conv0_b"], n_frames, O, C, 5, ACT_TANH, seg=fseg,
                           key="post_conv0")
        for l in (1, 2, 3):
            x = self.conv_gemm(x, w[f"post_conv{l}_w"], w[f"post_conv{l}_b"], n_frames, C, C, 5, ACT_TANH, seg=fseg,
                               key=f"post_conv{l}")
        final = torch.empty((n_frames, O), dtype=torch.float32, device=self.device)     # returned to the caller
        out = self.conv_gemm(x, w["post_conv4_w"], w["post_conv4_b"], n_frames, C, O, 5, ACT_NONE, seg=fseg,
                             residual=before, key="post_conv4", out=final)
        if chunks and chunk_cb is not None:
            for k, (u0, u1, f0, f1) in enumerate(chunks):          # not chunkable: hand everything over at the end
                chunk_cb(k, out, f0, f1)
        return out

    # ------------------------------------------------------------------ whole pass
    def upload(self, plan: BatchPlan):
        """All inputs of a batch in one pinned staging buffer (two alternate), one H2D copy on the upload stream.
        -> dict of device views + byte count."""
        P, B = plan.n_rows, plan.n_utts
        parts = [("ids", plan.ids.view(np.int32)), ("utt_off", plan.utt_off), ("row_utt", plan.row_utt),
                 ("row_phone", plan.row_phone), ("seg_lo", plan.seg_lo), ("seg_hi", plan.seg_hi)]
        if plan.dur is not None:
            if plan.dur.size and int(plan.dur.max()) > _lib.MAX_DURATION:
                raise ValueError(f"duration {int(plan.dur.max())} exceeds the supported maximum of {_lib.MAX_DURATION} frames "
                                 "per phoneme (FCL_MAX_DURATION; the reference's data cap is 50, preprocess.py:203)")
            parts.append(("dur", plan.dur.astype(np.int32, copy=False)))
        if plan.pitch is not None:
            parts.append(("pitch", plan.pitch.view(np.int32)))
            parts.append(("energy", plan.energy.view(np.int32)))
        offs, total = {}, 0
        for k, a in parts:
            offs[k] = (total, a.shape[0])
            total += (a.shape[0] + 3) // 4 * 4          # keep 16-byte alignment
        # Two pinned staging buffers and a dedicated upload stream: the H2D copy of batch i+1 runs while batch i computes
        # (on the launching stream it sat between two passes: ~0.1 ms of idle GPU per batch), and refilling a staging
        # buffer only waits for the copy that used it two uploads ago.
        slot = getattr(self, "_stage_slot", 0)
        self._stage_slot = slot ^ 1
        bufs = getattr(self, "_stage_bufs", None)
        if bufs is None:
            bufs = self._stage_bufs = [None, None]
            self._stage_evts = [None, None]
            self._up_stream = torch.cuda.Stream(device=self.device)
        if bufs[slot] is None or bufs[slot].numel() < total:
            bufs[slot] = torch.empty((max(total, 1 << 16),), dtype=torch.int32).pin_memory()   # cached: pinning is slow
            self._stage_evts[slot] = None
        if self._stage_evts[slot] is not None:
            self._stage_evts[slot].synchronize()        # the H2D copy that last read this buffer has finished
        stage = bufs[slot][:total]
        sn = stage.numpy()
        for k, a in parts:
            o, n = offs[k]
            sn[o:o + n] = a
        main = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self._up_stream):
            dev = stage.to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._up_stream)
        self._stage_evts[slot] = ev
        main.wait_event(ev)                             # the pass starts when its inputs have landed
        dev.record_stream(main)                         # allocated on the upload stream, used on the launching stream
        v = {k: dev[o:o + n] for k, (o, n) in offs.items()}
        v["ids"] = v["ids"].view(torch.int64)
        for k in ("pitch", "energy"):
            if k in v:
                v[k] = v[k].view(torch.float32)
        return v, total * 4

    class _Stage:
        def __init__(self, eng, name):
            self.eng, self.name = eng, name

        def __enter__(self):
            if self.eng.stage_events is not None:
                self.t0 = torch.cuda.Event(enable_timing=True)
                self.t0.record(torch.cuda.current_stream(self.eng.device))

        def __exit__(self, *a):
            if self.eng.stage_events is not None:
                t1 = torch.cuda.Event(enable_timing=True)
                t1.record(torch.cuda.current_stream(self.eng.device))
                self.eng.stage_events.append((self.name, self.t0, t1))

    def stage(self, name):
        return Engine._Stage(self, name)

    @torch.no_grad()
    def run(self, plan: BatchPlan, zoneout: float, dropout_p: float, dropout_seed: int,
            extras: bool = False, tile_rows=None, tf_y=None) -> BatchResult:
        """`tf_y` (F, odim) fp32 on the device, frames of all utterances in PROCESSING order: teacher-forced pass
        (the reference's forward(): decoder_sa.py:431-542); the predictors' outputs come back in `extras`."""
        with torch.cuda.device(self.device):
            d, h2d = self.upload(plan)
            return self.run_uploaded(plan, d, zoneout, dropout_p, dropout_seed, extras, tile_rows, h2d, tf_y=tf_y)

    @torch.no_grad()
    def run_uploaded(self, plan: BatchPlan, d: dict, zoneout: float, dropout_p: float, dropout_seed: int,
                     extras: bool = False, tile_rows=None, h2d: int = 0, out_chunks: int = 0, chunk_cb=None,
                     tf_y=None) -> BatchResult:
        """The pass proper, inputs already resident on the device (`d` from `upload`)."""
        self._arena_seq, self._in_pass = 0, not extras     # extras (tests) keep intermediates: no recycling
        try:
            with torch.cuda.device(self.device):           # launches go to the engine's device whatever is current outside
                return self._run_uploaded(plan, d, zoneout, dropout_p, dropout_seed, extras, tile_rows, h2d, out_chunks, chunk_cb,
                                          tf_y)
        finally:
            self._in_pass = False
            self._stream_handle = None

    def _run_uploaded(self, plan, d, zoneout, dropout_p, dropout_seed, extras, tile_rows, h2d, out_chunks=0, chunk_cb=None,
                      tf_y=None):
        hp = self.hp
        B, P = plan.n_utts, plan.n_rows
        ex = {"h2d_bytes": h2d}
        lens = np.diff(plan.utt_off.astype(np.int64))
        need_pred_dur = plan.dur is None
        main = torch.cuda.current_stream(self.device)
        self._stream_handle = main.cuda_stream
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(device=self.device)
        side = self._side

        def length_regulation(dur, F):
            """Integer work that depends only on the durations: scan / sort, frame map, postnet tiles, decoder tile
            schedule. With forced durations it runs on a side stream concurrently with the encoder and predictors."""
            with self.stage("len_reg"):
                frame_off, utt_frame_off, order, totals = self.len_reg_scan(dur, d["utt_off"], B)
            sched = self.decoder_schedule(order, dur, P) if self.precision != "fp32" and self.bf16_decoder else None
            fmap = pos = ftiles = None
            if F is not None and need_fmap:
                with self.stage("frame_map"):
                    fmap, pos = self.frame_map(frame_off, utt_frame_off, P, B, F, want_position=extras)
                    ftiles = self.conv_tiles(utt_frame_off, B, (F + 127) // 128 + B)
            return frame_off, utt_frame_off, order, totals, sched, fmap, pos, ftiles

        # the frame -> (row, step) map and the per-utterance frame tiles only serve the layer-by-layer / fused-stack postnet
        # (and the diagnostics of `extras`); the image postnet works from the frame offsets alone
        need_fmap = extras or tf_y is not None or not (self.precision != "fp32" and self.use_img_postnet)
        if tf_y is not None and need_pred_dur:
            raise ValueError("teacher forcing needs the ground-truth durations (forward(): extras / ds)")
        lr = None
        if not need_pred_dur:
            if not self.skip_zero_durations and (plan.dur == 0).any():
                raise ValueError("zero durations are outside the reference's working domain "
                                 "(nets/modules/decoder_sa.py:575); skip_zero_durations=True drops those phonemes from "
                                 "the decoder the way the reference's forward() does (decoder_sa.py:459-463)")
            per_utt = np.add.reduceat(plan.dur.astype(np.int64), plan.utt_off[:-1].astype(np.int64))
            ufo = np.concatenate([[0], np.cumsum(per_utt)])
            F = int(ufo[-1])
            use_side = P >= 4096                           # tiny batches: the cross-stream hand-over costs more than it hides
            if use_side:
                side.wait_stream(main)                     # inputs uploaded, previous pass finished with the arena
                with torch.cuda.stream(side):
                    self._stream_handle = side.cuda_stream
                    lr = length_regulation(d["dur"], F)
                self._stream_handle = main.cuda_stream
            else:
                lr = length_regulation(d["dur"], F)
        with self.stage("encoder"):
            if self.precision != "fp32" and self.use_img_convs and self.blstm_whh_bf16 is not None:
                seg = (d["seg_lo"], d["seg_hi"], None, self.pad_rows(d["utt_off"], B, P))
            else:
                seg = (d["seg_lo"], d["seg_hi"], self.conv_tiles(d["utt_off"], B, int(((lens + 127) // 128).sum())))
            h = self.encoder(d["ids"], d["utt_off"], seg, B, lens=lens)
        dlog = dur_pred = None
        with self.stage("predictors"):
            if need_pred_dur or extras or tf_y is not None:
                dlog, dur_pred = self.predictor("dur", h, seg, want_dur=True)
            dur = dur_pred if need_pred_dur else d["dur"]
            pitch_pred = energy_pred = None
            if plan.pitch is None or tf_y is not None:
                pitch_pred, _ = self.predictor("pitch", h, seg)
                energy_pred, _ = self.predictor("energy", h, seg)
            if plan.pitch is None:
                pitch, energy = pitch_pred, energy_pred
            else:
                pitch, energy = d["pitch"], d["energy"]
        if need_pred_dur:
            frame_off, utt_frame_off, order, totals, sched, _, _, _ = length_regulation(dur, None)
            host = torch.cat([totals, utt_frame_off]).cpu().numpy()      # the one data-dependent D2H sync
            F, ufo = int(host[0]), host[2:].astype(np.int64)
            if int(host[1]) >= _lib.MAX_DURATION:
                raise ValueError(f"a predicted duration reached the supported maximum of {_lib.MAX_DURATION} frames per "
                                 "phoneme (FCL_MAX_DURATION): the reference has no cap, refusing to truncate silently")
            if not self.skip_zero_durations and int((dur == 0).sum()) != 0:
                raise ValueError("predicted zero durations: outside the reference's working domain "
                                 "(nets/modules/decoder_sa.py:575); pass dur= or skip_zero_durations=True")
            if F == 0:
                raise ValueError("every predicted duration is zero: no frames to decode")
            fmap = pos = ftiles = None
            if need_fmap:
                with self.stage("frame_map"):
                    fmap, pos = self.frame_map(frame_off, utt_frame_off, P, B, F, want_position=extras)
                    ftiles = self.conv_tiles(utt_frame_off, B, (F + 127) // 128 + B)
        else:
            frame_off, utt_frame_off, order, totals, sched, fmap, pos, ftiles = lr
            if use_side:
                main.wait_stream(side)
        tc_dec = self.precision != "fp32" and self.bf16_decoder
        with self.stage("embed_add"):       # pitch / energy embeddings + add; on the tensor-core path straight into the
            hn, hn_img = self.embed_add(h, pitch, energy, seg, order=order if tc_dec else None,   # decoder's operand image
                                        want_rows=extras or not tc_dec)
        if tf_y is not None and tuple(tf_y.shape) != (F, hp.odim):
            raise ValueError(f"teacher forcing: {tuple(tf_y.shape)} target frames for {F} frames of duration")
        before = self.decoder(hn, dur, frame_off, order, d["row_utt"], d["row_phone"], F, zoneout, dropout_p,
                              dropout_seed, tile_rows, schedule=sched,
                              tf=(tf_y, fmap[0], fmap[1]) if tf_y is not None else None, hn_img=hn_img)
        with self.stage("postnet"):
            chunks = output_chunks(ufo, out_chunks) if (out_chunks and chunk_cb is not None) else None
            out = self.postnet(before, (fmap[2] if fmap is not None else None, fmap[3] if fmap is not None else None, ftiles,
                                        (utt_frame_off, B)), F, chunks, chunk_cb)
        if tf_y is not None:
            ex.update(before=before.clone(), dlog=dlog.clone(), pitch_pred=pitch_pred.clone(), energy_pred=energy_pred.clone())
        if extras:
            ex.update(h=h, dlog=dlog, dur_pred=dur_pred, pitch=pitch, energy=energy, hn=hn, before=before,
                      frame_off=frame_off, order=order, frame_row=fmap[0], frame_step=fmap[1], position=pos,
                      totals=totals, utt_frame_off_dev=utt_frame_off)
        return BatchResult(out=out, utt_frame_off=ufo, perm=plan.perm, extras=ex)
