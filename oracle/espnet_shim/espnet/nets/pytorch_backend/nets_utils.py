"""Stub of the three espnet helpers the reference nets import
(reference: nets/teacher_training/e2e_tts_tacotron2_sa.py:12, nets/modules/decoder_sa.py:12,16).

Test infrastructure only -- written from the documented espnet 0.8 behaviour,
not copied; espnet is not installed in this image.
"""
import torch


def pad_list(xs, pad_value):
    """Stack variable-length tensors into (B, Tmax, ...) filled with pad_value."""
    n = len(xs)
    tmax = max(x.size(0) for x in xs)
    out = xs[0].new_full((n, tmax) + tuple(xs[0].shape[1:]), pad_value)
    for i, x in enumerate(xs):
        out[i, : x.size(0)] = x
    return out


def make_pad_mask(lengths, xs=None, length_dim=-1):
    """Bool mask (B, Tmax), True at padded positions."""
    if not isinstance(lengths, list):
        lengths = lengths.tolist()
    tmax = int(max(lengths)) if xs is None else xs.size(length_dim)
    ar = torch.arange(tmax, dtype=torch.int64).unsqueeze(0)
    mask = ar >= torch.tensor(lengths, dtype=torch.int64).unsqueeze(1)
    if xs is not None:
        mask = mask.to(xs.device)
    return mask


def make_non_pad_mask(lengths, xs=None, length_dim=-1):
    return ~make_pad_mask(lengths, xs, length_dim)
