"""Seeded synthetic weights and inputs.

`state_dict_spec` lists every tensor of the reference checkpoint layout
(produced by nets/teacher_training/e2e_tts_tacotron2_sa.py:365-468
and the nets/modules constructors) so the same seeded weights can be loaded
into the reference model (in the build container) and into the B200 model
(anywhere). There is no network: checkpoints do not exist, weights are random.
"""
from __future__ import annotations

import hashlib
from collections import OrderedDict

import numpy as np
import torch

from .hparams import HParams


def state_dict_spec(hp: HParams, student_kd_keys: bool = False, teacher: HParams | None = None):
    """-> OrderedDict name -> (shape, kind). kind in
    {xavier_relu, xavier_tanh, uniform_fan, bn_weight, bn_bias, bn_mean, bn_var, bn_count,
     ln_weight, ln_bias, normal}."""
    E, H, C, U, O = hp.eunits, hp.dunits, hp.postnet_chans, hp.prenet_units, hp.odim
    PC, PK = hp.predictor_chans, hp.predictor_kernel
    s = OrderedDict()
    s["enc.embed.weight"] = ((hp.idim, hp.embed_dim), "embed")
    for l in range(hp.econv_layers):
        cin = hp.embed_dim if l == 0 else hp.econv_chans
        s[f"enc.convs.{l}.0.weight"] = ((hp.econv_chans, cin, hp.econv_filts), "xavier_relu")
        _bn(s, f"enc.convs.{l}.1", hp.econv_chans)
    hd = E // 2
    for suf in ("", "_reverse"):
        s[f"enc.blstm.weight_ih_l0{suf}"] = ((4 * hd, hp.econv_chans), ("uniform_fan", hd))
        s[f"enc.blstm.weight_hh_l0{suf}"] = ((4 * hd, hd), ("uniform_fan", hd))
        s[f"enc.blstm.bias_ih_l0{suf}"] = ((4 * hd,), ("uniform_fan", hd))
        s[f"enc.blstm.bias_hh_l0{suf}"] = ((4 * hd,), ("uniform_fan", hd))
    if student_kd_keys:
        tch = teacher
        s["enc.embed_proj.weight"] = ((tch.embed_dim, hp.embed_dim), ("uniform_fan", hp.embed_dim))
        s["enc.convs_proj.0.weight"] = ((tch.econv_chans, hp.econv_chans), ("uniform_fan", hp.econv_chans))
        s["enc.blstm_proj.weight"] = ((tch.eunits, E), ("uniform_fan", E))
    in0 = E + U + (1 if hp.append_position else 0)
    s["dec.lstm.0.cell.weight_ih"] = ((4 * H, in0), ("uniform_fan", H))
    s["dec.lstm.0.cell.weight_hh"] = ((4 * H, H), ("uniform_fan", H))
    s["dec.lstm.0.cell.bias_ih"] = ((4 * H,), ("uniform_fan", H))
    s["dec.lstm.0.cell.bias_hh"] = ((4 * H,), ("uniform_fan", H))
    s["dec.lstm.1.cell.weight_ih"] = ((4 * H, H), ("uniform_fan", H))
    s["dec.lstm.1.cell.weight_hh"] = ((4 * H, H), ("uniform_fan", H))
    s["dec.lstm.1.cell.bias_ih"] = ((4 * H,), ("uniform_fan", H))
    s["dec.lstm.1.cell.bias_hh"] = ((4 * H,), ("uniform_fan", H))
    s["dec.prenet.prenet.0.0.weight"] = ((U, O), ("uniform_fan", O))
    s["dec.prenet.prenet.0.0.bias"] = ((U,), ("uniform_fan", O))
    s["dec.prenet.prenet.1.0.weight"] = ((U, U), ("uniform_fan", U))
    s["dec.prenet.prenet.1.0.bias"] = ((U,), ("uniform_fan", U))
    for l in range(hp.postnet_layers):
        cin = O if l == 0 else C
        cout = O if l == hp.postnet_layers - 1 else C
        s[f"dec.postnet.postnet.{l}.0.weight"] = ((cout, cin, hp.postnet_filts), "xavier_tanh")
        _bn(s, f"dec.postnet.postnet.{l}.1", cout)
    s["dec.feat_out.weight"] = ((O, H + E), ("uniform_fan", H + E))
    if student_kd_keys:
        tch = teacher
        s["dec.prenet_proj.weight"] = ((tch.prenet_units, U), ("uniform_fan", U))
        s["dec.lstm_proj.weight"] = ((tch.dunits, H), ("uniform_fan", H))
        s["dec.post_proj.weight"] = ((tch.postnet_chans, C), ("uniform_fan", C))
    for name in ("duration_predictor", "pitch_predictor", "energy_predictor"):
        s[f"{name}.conv.0.0.weight"] = ((PC, E, PK), ("uniform_fan", E * PK))
        s[f"{name}.conv.0.0.bias"] = ((PC,), ("uniform_fan", E * PK))
        s[f"{name}.conv.0.2.weight"] = ((PC,), "ln_weight")
        s[f"{name}.conv.0.2.bias"] = ((PC,), "ln_bias")
        s[f"{name}.conv.1.0.weight"] = ((PC, PC, PK), ("uniform_fan", PC * PK))
        s[f"{name}.conv.1.0.bias"] = ((PC,), ("uniform_fan", PC * PK))
        s[f"{name}.conv.1.2.weight"] = ((PC,), "ln_weight")
        s[f"{name}.conv.1.2.bias"] = ((PC,), "ln_bias")
        s[f"{name}.linear.weight"] = ((1, PC), ("uniform_fan", PC))
        s[f"{name}.linear.bias"] = ((1,), ("uniform_fan", PC))
    for name in ("pitch_embed", "energy_embed"):
        s[f"{name}.0.weight"] = ((E, 1, hp.embed_kernel), ("uniform_fan", hp.embed_kernel))
        s[f"{name}.0.bias"] = ((E,), ("uniform_fan", hp.embed_kernel))
    if student_kd_keys:
        s["pemb_proj.weight"] = ((teacher.eunits, E), ("uniform_fan", E))
        s["eemb_proj.weight"] = ((teacher.eunits, E), ("uniform_fan", E))
    return s


def _bn(s, prefix, c):
    s[f"{prefix}.weight"] = ((c,), "bn_weight")
    s[f"{prefix}.bias"] = ((c,), "bn_bias")
    s[f"{prefix}.running_mean"] = ((c,), "bn_mean")
    s[f"{prefix}.running_var"] = ((c,), "bn_var")
    s[f"{prefix}.num_batches_tracked"] = ((), "bn_count")


def random_state_dict(hp: HParams, seed: int, student_kd_keys: bool = False,
                      teacher: HParams | None = None) -> "OrderedDict[str, torch.Tensor]":
    """Seeded random-init weights with the reference's init distributions
    (xavier-uniform convs: encoder_sa.py:15-18 / decoder_sa.py:20-23; torch
    defaults elsewhere) and *randomised* BatchNorm/LayerNorm statistics
    (identity BN would hide folding bugs)."""
    g = torch.Generator().manual_seed(1000003 * seed + 17)
    sd = OrderedDict()

    def u(shape, bound):
        return (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound

    for name, (shape, kind) in state_dict_spec(hp, student_kd_keys, teacher).items():
        if isinstance(kind, tuple):
            fan = kind[1]
            t = u(shape, 1.0 / np.sqrt(fan))
        elif kind == "embed":
            t = torch.randn(shape, generator=g, dtype=torch.float32)
            t[0].zero_()                                  # padding_idx=0 row
        elif kind in ("xavier_relu", "xavier_tanh"):
            gain = np.sqrt(2.0) if kind == "xavier_relu" else 5.0 / 3.0
            fan_in, fan_out = shape[1] * shape[2], shape[0] * shape[2]
            t = u(shape, gain * np.sqrt(6.0 / (fan_in + fan_out)))
        elif kind in ("bn_weight", "ln_weight"):
            t = torch.rand(shape, generator=g) + 0.5      # U(0.5,1.5)
        elif kind in ("bn_bias", "ln_bias"):
            t = torch.randn(shape, generator=g) * 0.1
        elif kind == "bn_mean":
            t = torch.randn(shape, generator=g) * 0.2
        elif kind == "bn_var":
            t = torch.rand(shape, generator=g) + 0.5
        elif kind == "bn_count":
            t = torch.tensor(100, dtype=torch.int64)
        else:
            raise KeyError(kind)
        sd[name] = t
    return sd


def state_dict_digest(sd) -> str:
    """sha256 over all float tensors, to detect RNG drift between torch builds."""
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


# ----------------------------------------------------------------------------- inputs

def utterance_lengths(batch: int, seed: int, fixed: int | None = None) -> np.ndarray:
    """LJSpeech-shaped phoneme counts: clip(round(N(80,25)),10,150)."""
    if fixed is not None:
        return np.full(batch, fixed, dtype=np.int64)
    r = np.random.RandomState(seed * 7919 + 1)
    return np.clip(np.rint(r.normal(80.0, 25.0, size=batch)), 10, 150).astype(np.int64)


def phoneme_ids(n: int, idim: int, rs: np.random.RandomState) -> np.ndarray:
    return rs.randint(1, idim, size=n).astype(np.int64)            # 0 = PAD never appears


def durations_ljspeech(n: int, rs: np.random.RandomState) -> np.ndarray:
    """clip(1 + Poisson(6), 1, 50): mean ~7 frames per phoneme."""
    return np.clip(1 + rs.poisson(6.0, size=n), 1, 50).astype(np.int64)


def durations_stress(n: int, rs: np.random.RandomState) -> np.ndarray:
    """clip(round(LogNormal(1.5,0.8)),1,40) with at least one d=40."""
    d = np.clip(np.rint(rs.lognormal(1.5, 0.8, size=n)), 1, 40).astype(np.int64)
    d[rs.randint(0, n)] = 40
    return d


def synth_batch(batch: int, seed: int, idim: int = 76, fixed_len: int | None = None,
                stress: bool = False):
    """-> (list of id arrays, list of duration arrays)."""
    rs = np.random.RandomState(seed * 104729 + 3)
    lens = utterance_lengths(batch, seed, fixed_len)
    xs, ds = [], []
    for n in lens:
        xs.append(phoneme_ids(int(n), idim, rs))
        ds.append(durations_stress(int(n), rs) if stress else durations_ljspeech(int(n), rs))
    return xs, ds
