"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

from fcl_taco2_b200 import hparams, synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    g = {k: z[k] for k in z.files}
    for k in ("kind", "weights_sha256"):
        g[k] = str(g[k])
    for k in ("weight_seed", "dropout_seed", "utt_index"):
        g[k] = int(g[k])
    g["dropout_rate"] = float(g["dropout_rate"])
    return g


_SD = {}


def weights(kind, seed):
    key = (kind, seed)
    if key not in _SD:
        _SD[key] = synth.random_state_dict(hparams.preset(kind), seed, kind == "S", hparams.preset("T"))
    return _SD[key]


def err(a, b):
    a = torch.as_tensor(a, dtype=torch.float32)
    b = torch.as_tensor(b, dtype=torch.float32)
    return float((a - b).abs().max()), float((a - b).abs().mean())
