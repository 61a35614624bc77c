"""One-time weight repacking: reference `state_dict` -> kernel layouts.

All transforms are exact re-arrangements except the eval-mode BatchNorm fold
(scale into the conv weight, shift into a bias; encoder_sa.py:74, decoder_sa.py:213),
done in fp32 on the host. Gate-interleaving puts i,f,g,o of one hidden unit in 4
adjacent columns (torch.nn.LSTM/LSTMCell order is [i;f;g;o] blocks of H rows), so the
LSTM cell update never leaves registers.
"""
from __future__ import annotations

import torch

from .hparams import HParams

BN_EPS = 1e-5


def op_dtype():
    """torch dtype of the 16-bit tensor-core operands (weights are packed in the format the library was built for:
    fp16 by default, bf16 with -DFCL_OPERANDS_BF16; csrc/umma.cuh)."""
    from . import _lib
    return torch.bfloat16 if _lib.operand_format() == "bf16" else torch.float16


def _to_op(x: torch.Tensor) -> torch.Tensor:
    dt = op_dtype()
    if dt is torch.float16:
        x = x.clamp(-65504.0, 65504.0)          # saturate like the kernels' conversions
    return x.to(dt)


def _interleave_gates(w: torch.Tensor, hidden: int) -> torch.Tensor:
    """(4H, K) rows [i;f;g;o] -> (K, 4H) with column u*4+g."""
    k = w.shape[1]
    return w.view(4, hidden, k).permute(2, 1, 0).reshape(k, 4 * hidden).contiguous()


def _interleave_bias(b: torch.Tensor, hidden: int) -> torch.Tensor:
    return b.view(4, hidden).t().reshape(4 * hidden).contiguous()


def _fold_bn_conv(sd, conv_key: str, bn_prefix: str):
    """Conv1d weight (Cout, Cin, k) + BatchNorm1d eval -> packed (k, Cin, Cout), bias (Cout)."""
    w = sd[conv_key].float()
    scale = sd[bn_prefix + ".weight"].float() / torch.sqrt(sd[bn_prefix + ".running_var"].float() + BN_EPS)
    shift = sd[bn_prefix + ".bias"].float() - sd[bn_prefix + ".running_mean"].float() * scale
    w = w * scale.view(-1, 1, 1)
    return w.permute(2, 1, 0).contiguous(), shift.contiguous()


def pack_fp32(sd, hp: HParams) -> dict:
    """-> dict of contiguous fp32 CPU tensors in the layouts include/fcl_taco2.h documents."""
    E, H, U, O = hp.eunits, hp.dunits, hp.prenet_units, hp.odim
    hd = E // 2
    f = lambda k: sd[k].detach().float().cpu()
    out = {"embed": f("enc.embed.weight").contiguous()}
    for l in range(3):
        out[f"enc_conv{l}_w"], out[f"enc_conv{l}_b"] = _fold_bn_conv(
            {k: f(k) for k in sd if k.startswith(f"enc.convs.{l}.")}, f"enc.convs.{l}.0.weight", f"enc.convs.{l}.1")
    # BiLSTM: input projection for both directions as one (1, E, 8*hd) "conv", recurrent (2, hd, 4*hd)
    wih = [_interleave_gates(f("enc.blstm.weight_ih_l0" + s), hd) for s in ("", "_reverse")]
    bih = [_interleave_bias(f("enc.blstm.bias_ih_l0" + s) + f("enc.blstm.bias_hh_l0" + s), hd) for s in ("", "_reverse")]
    out["blstm_wih"] = torch.cat(wih, dim=1).unsqueeze(0).contiguous()
    out["blstm_b"] = torch.cat(bih).contiguous()
    out["blstm_whh"] = torch.stack([_interleave_gates(f("enc.blstm.weight_hh_l0" + s), hd) for s in ("", "_reverse")]).contiguous()
    for name, short in (("duration_predictor", "dur"), ("pitch_predictor", "pitch"), ("energy_predictor", "energy")):
        for l in range(2):
            out[f"{short}_conv{l}_w"] = f(f"{name}.conv.{l}.0.weight").permute(2, 1, 0).contiguous()
            out[f"{short}_conv{l}_b"] = f(f"{name}.conv.{l}.0.bias").contiguous()
            out[f"{short}_ln{l}_g"] = f(f"{name}.conv.{l}.2.weight").contiguous()
            out[f"{short}_ln{l}_b"] = f(f"{name}.conv.{l}.2.bias").contiguous()
        out[f"{short}_head_w"] = f(f"{name}.linear.weight").reshape(-1).contiguous()
        out[f"{short}_head_b"] = f(f"{name}.linear.bias").reshape(-1).contiguous()
    for name, short in (("pitch_embed", "pemb"), ("energy_embed", "eemb")):
        out[f"{short}_w"] = f(f"{name}.0.weight").reshape(E, -1).contiguous()
        out[f"{short}_b"] = f(f"{name}.0.bias").contiguous()
    # decoder
    wih0, whh0 = f("dec.lstm.0.cell.weight_ih"), f("dec.lstm.0.cell.weight_hh")
    assert wih0.shape[1] == E + U + 1
    out["dec_g0h_w"] = _interleave_gates(wih0[:, :E].contiguous(), H).unsqueeze(0).contiguous()         # (1, E, 4H)
    out["dec_g0h_b"] = _interleave_bias(f("dec.lstm.0.cell.bias_ih") + f("dec.lstm.0.cell.bias_hh"), H)
    out["dec_w0"] = torch.cat([_interleave_gates(wih0[:, E:E + U].contiguous(), H), _interleave_gates(whh0, H)], dim=0).contiguous()
    out["dec_wpos"] = _interleave_bias(wih0[:, E + U].contiguous(), H)
    out["dec_w1"] = torch.cat([_interleave_gates(f("dec.lstm.1.cell.weight_ih"), H),
                               _interleave_gates(f("dec.lstm.1.cell.weight_hh"), H)], dim=0).contiguous()
    out["dec_b1"] = _interleave_bias(f("dec.lstm.1.cell.bias_ih") + f("dec.lstm.1.cell.bias_hh"), H)
    out["dec_wp0"] = f("dec.prenet.prenet.0.0.weight").t().contiguous()
    out["dec_bp0"] = f("dec.prenet.prenet.0.0.bias").contiguous()
    out["dec_wp1"] = f("dec.prenet.prenet.1.0.weight").t().contiguous()
    out["dec_bp1"] = f("dec.prenet.prenet.1.0.bias").contiguous()
    wfeat = f("dec.feat_out.weight")
    assert wfeat.shape == (O, H + E)
    out["dec_wf"] = wfeat[:, :H].t().contiguous()                                                        # (H, O)
    out["dec_y0h_w"] = wfeat[:, H:].t().contiguous().unsqueeze(0).contiguous()                           # (1, E, O)
    for l in range(5):
        out[f"post_conv{l}_w"], out[f"post_conv{l}_b"] = _fold_bn_conv(
            {k: f(k) for k in sd if k.startswith(f"dec.postnet.postnet.{l}.")},
            f"dec.postnet.postnet.{l}.0.weight", f"dec.postnet.postnet.{l}.1")
    return out


# ----------------------------------------------------------------------------- bf16 tensor-core packing
def choose_ntile(cout: int) -> int:
    """Largest multiple of 16 that divides cout and is <= 256."""
    for nt in range(256, 15, -16):
        if cout % nt == 0:
            return nt
    raise ValueError(f"cout={cout} has no tile width that is a multiple of 16")


def choose_kstage(cin: int) -> int:
    for ks in (64, 80, 48, 32, 16):
        if cin % ks == 0:
            return ks
    raise ValueError(f"cin={cin} must be a multiple of 16")


def pack_conv_bf16(w: torch.Tensor, ntile: int | None = None, kstage: int | None = None):
    """(taps, cin, cout) fp32 -> bf16 blocks [cout/ntile][cin/kstage][taps][kstage/8][ntile][8]
    (UMMA K-major no-swizzle core-matrix image of each (ntile x kstage) B stage; csrc/umma.cuh).
    -> (flat bf16 tensor, ntile, kstage)"""
    taps, cin, cout = w.shape
    ntile = ntile or choose_ntile(cout)
    kstage = kstage or choose_kstage(cin)
    assert cout % ntile == 0 and cin % kstage == 0 and kstage % 16 == 0
    x = w.reshape(taps, cin // kstage, kstage // 8, 8, cout // ntile, ntile)      # t, kc, k8, j, nt, n
    x = x.permute(4, 1, 0, 2, 5, 3).contiguous()                                   # nt, kc, t, k8, n, j
    return _to_op(x).reshape(-1).contiguous(), ntile, kstage


def choose_nb(cout: int) -> int:
    """MMA N of the image convolutions (csrc/conv_img_bf16.cu): largest multiple of 64 <= 256 dividing cout."""
    for nb in (256, 192, 128, 64):
        if cout % nb == 0:
            return nb
    raise ValueError(f"cout={cout} must be a multiple of 64")


def pack_conv_pair(w: torch.Tensor, nb: int | None = None):
    """(taps, cin, cout) fp32 -> bf16 half-stage blocks [cout/nb][cin/64][taps][2 halves][8][nb/2][8] for the
    cta_group::2 image convolutions: each CTA of a pair loads its half (nb/2 output columns x 64 k) of every weight
    stage as one contiguous block, already in the UMMA K-major no-swizzle core-matrix order. -> (flat bf16, nb)"""
    taps, cin, cout = w.shape
    nb = nb or choose_nb(cout)
    assert cin % 64 == 0 and cout % nb == 0 and nb % 64 == 0 and nb <= 256
    x = w.reshape(taps, cin // 64, 8, 8, cout // nb, 2, nb // 2)                   # t, kc, k8, j, blk, half, n
    x = x.permute(4, 1, 0, 5, 2, 6, 3).contiguous()                                 # blk, kc, t, half, k8, n, j
    return _to_op(x).reshape(-1).contiguous(), nb


IMG_KEYS = ["enc_conv0", "enc_conv1", "enc_conv2", "blstm_wih", "dur_conv0", "dur_conv1", "pitch_conv0", "pitch_conv1",
            "energy_conv0", "energy_conv1"]


def pack_img(packed_fp32: dict) -> dict:
    """Pair-split bf16 weights of the front-end image convolutions, or {} when a shape does not fit the kernel."""
    out = {}
    try:
        for key in IMG_KEYS:
            w = packed_fp32[key if key == "blstm_wih" else key + "_w"]
            if w.shape[1] % 64 or w.shape[1] > 512 or w.shape[2] % 64 or w.shape[2] > 2048:
                return {}
            out[key] = pack_conv_pair(w)
    except (ValueError, AssertionError):
        return {}
    return out


def _pad_to(x: torch.Tensor, dim: int, mult: int) -> torch.Tensor:
    n = x.shape[dim]
    m = (n + mult - 1) // mult * mult
    if m == n:
        return x
    shape = list(x.shape)
    shape[dim] = m
    out = torch.zeros(shape, dtype=x.dtype)
    out.narrow(dim, 0, n).copy_(x)
    return out


def pack_img_postnet(packed_fp32: dict) -> dict:
    """Postnet layers (decoder_sa.py:176-263: Conv1d k5 no bias + BatchNorm eval, folded) for the image convolutions:
    input / output channels zero-padded to multiples of 64 (80 -> 128 mel channels). -> {key: (weights, nb, cin, cout, bias)}
    or {} when a layer does not fit the kernel."""
    out = {}
    try:
        for l in range(5):
            w = _pad_to(_pad_to(packed_fp32[f"post_conv{l}_w"], 1, 64), 2, 64)
            if w.shape[1] > 512 or w.shape[2] > 2048 or w.shape[0] != 5:
                return {}
            wp, nb = pack_conv_pair(w)
            out[f"post_conv{l}"] = (wp, nb, w.shape[1], w.shape[2], _pad_to(packed_fp32[f"post_conv{l}_b"], 0, 64).contiguous())
    except (ValueError, AssertionError):
        return {}
    return out


STACK_KSTAGE = 32

GEMM_KEYS = ["enc_conv0", "enc_conv1", "enc_conv2", "blstm_wih", "dur_conv0", "dur_conv1", "pitch_conv0", "pitch_conv1",
             "energy_conv0", "energy_conv1", "post_conv0", "post_conv1", "post_conv2",
             "post_conv3", "post_conv4"]


def pack_bf16(packed_fp32: dict) -> dict:
    """bf16 tensor-core operands for every conv/linear GEMM, from the fp32 packed dict."""
    out = {}
    for key in GEMM_KEYS:
        w = packed_fp32[key if key == "blstm_wih" else key + "_w"]
        out[key] = pack_conv_bf16(w)
    # postnet layers for the fused stack: K stages of STACK_KSTAGE channels (first layer's input channels zero-padded
    # to a multiple of it) so the weight-ring slots stay small and several CTAs share an SM
    for l in range(5):
        w = packed_fp32[f"post_conv{l}_w"]
        taps, cin, cout = w.shape
        if cin % STACK_KSTAGE:
            pad = torch.zeros(taps, (cin + STACK_KSTAGE - 1) // STACK_KSTAGE * STACK_KSTAGE, cout)
            pad[:, :cin] = w
            w = pad
        if cout <= 256 and cout % 16 == 0:
            out[f"post_stack{l}"] = pack_conv_bf16(w, cout, STACK_KSTAGE)
    return out


def _split_halves(blocks: torch.Tensor, n: int) -> torch.Tensor:
    """bf16 stage blocks [..][8][n][8] -> [..][2 halves][8][n/2][8]: each CTA of a pair loads its half of the columns."""
    b = blocks.reshape(-1, 8, 2, n // 2, 8).permute(0, 2, 1, 3, 4).contiguous()
    return b.reshape(-1)


def pack_decoder_stream(packed_fp32: dict, hp, pair: bool = False) -> torch.Tensor:
    """bf16 weight stream of the tensor-core decoder, in the order the kernel consumes it each step
    (csrc/decoder_bf16.cu):
        P1  prenet.1                                   rows [x1]            256 cols
        L0  cell 0 gates (gate-interleaved)            rows [h ; z0 ; x2]   4H cols
        L1  cell 1 gates                               rows [z1 ; z0']      4H cols
        F   feat_out                                   rows [h ; z1']       odim cols
        PC  prenet.0 composed with feat_out            rows [h ; z1']       256 cols
    PC = W_feat^T W_p0^T: feat_out has no activation, so prenet.0(y) = relu([z1'|h] PC + b_p0) (fp32 product,
    rounded to bf16 once). Every block is the UMMA core-matrix image of one (256-or-odim columns x 64 k) B stage."""
    U, H, O, E = hp.prenet_units, hp.dunits, hp.odim, hp.eunits
    assert U == 256 and H % 64 == 0 and E % 64 == 0 and O % 16 == 0 and O <= 128
    w0 = packed_fp32["dec_w0"]                       # rows [prenet part (U) ; W_hh0 (H)], cols 4H gate-interleaved
    w0h = packed_fp32["dec_g0h_w"][0]                # (E, 4H): h part of W_ih0
    l0 = torch.cat([w0h, w0[U:], w0[:U]], dim=0)     # [h ; z0 ; x2]
    w1 = packed_fp32["dec_w1"]                       # rows [W_ih1 (z0') ; W_hh1 (z1)]
    l1 = torch.cat([w1[H:], w1[:H]], dim=0)          # [z1 ; z0']
    feat = torch.cat([packed_fp32["dec_y0h_w"][0], packed_fp32["dec_wf"]], dim=0)      # (E + H, O) rows [h ; z1']
    pc = (feat.double() @ packed_fp32["dec_wp0"].double()).float()                      # (E + H, U)
    if pair:
        # cta_group::2 kernel: every stage block is split along N into the two CTAs' halves; feat_out's odim columns are
        # zero-padded to 128 so that both halves are whole core-matrix rows
        featp = torch.zeros(feat.shape[0], 128)
        featp[:, :O] = feat
        wide = [packed_fp32["dec_wp1"], l0, l1, pc]
        pk = [_split_halves(pack_conv_bf16(w.unsqueeze(0), 256, 64)[0], 256) for w in wide]
        pf = _split_halves(pack_conv_bf16(featp.unsqueeze(0), 128, 64)[0], 128)
        return torch.cat([pk[0], pk[1], pk[2], pf, pk[3]]).contiguous()
    parts = [
        pack_conv_bf16(packed_fp32["dec_wp1"].unsqueeze(0), 256, 64)[0],
        pack_conv_bf16(l0.unsqueeze(0), 256, 64)[0],
        pack_conv_bf16(l1.unsqueeze(0), 256, 64)[0],
        pack_conv_bf16(feat.unsqueeze(0), O, 64)[0],
        pack_conv_bf16(pc.unsqueeze(0), 256, 64)[0],
    ]
    return torch.cat(parts).contiguous()


def pack_bilstm_whh_bf16(packed_fp32: dict) -> torch.Tensor:
    """W_hh of both directions as bf16 UMMA B stages: [dir][4H/256 chunks][H/64 k-stages][8][256][8]
    (csrc/bilstm_bf16.cu). Input: packed_fp32["blstm_whh"] (2, H, 4H) gate-interleaved columns."""
    whh = packed_fp32["blstm_whh"]
    return torch.cat([pack_conv_bf16(whh[d].unsqueeze(0), 256, 64)[0] for d in range(2)]).contiguous()
