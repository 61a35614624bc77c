"""Import the UNMODIFIED reference modules (build container only).

TEST INFRASTRUCTURE. Only tests/, oracle/make_golden.py and the validation of
oracle/restate.py use this; it needs /root/reference, which does not exist on
the GPU box. Nothing under fcl_taco2_b200/ imports it.

The reference imports `espnet` (not installed): oracle/espnet_shim provides the
8 symbols it needs. `nets/` has no __init__.py: namespace
packages resolve with /root/reference on sys.path.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import os
import sys

import torch

REFERENCE_ROOT = os.environ.get("FCL_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "espnet_shim")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "nets"))


def _paths():
    for p in (_SHIM, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)


def _yaml_ns(name):
    import yaml
    with open(os.path.join(REFERENCE_ROOT, "conf", name)) as f:
        y = yaml.safe_load(f)
    ns = argparse.Namespace(**{k.replace("-", "_"): v for k, v in y.items()})
    ns.encoder_resume = None
    return ns


def com_args():
    return argparse.Namespace(
        use_fe_condition=True, append_position=True,
        distill_output_knowledge=True, distill_encoder_knowledge=True,
        distill_decoder_knowledge=True, distill_prosody_knowledge=True,
        is_train=True, share_proj=True)


def build(kind: str, idim: int = 76, odim: int = 80):
    """kind 'T' -> nets.teacher_training...Tacotron2_sa from conf/...sa.yaml;
    kind 'S' -> nets.knowledge_distillation...kd_student.Tacotron2_sa from
    conf/...sa.student.yaml (+ teacher yaml for the KD projection sizes)."""
    _paths()
    with contextlib.redirect_stdout(io.StringIO()):
        if kind == "T":
            from nets.teacher_training.e2e_tts_tacotron2_sa import Tacotron2_sa
            m = Tacotron2_sa(idim, odim, _yaml_ns("train_pytorch_tacotron2.sa.yaml"), com_args())
        elif kind == "S":
            from nets.knowledge_distillation.e2e_tts_tacotron2_sa_kd_student import Tacotron2_sa
            m = Tacotron2_sa(idim, odim, _yaml_ns("train_pytorch_tacotron2.sa.student.yaml"),
                             com_args(), _yaml_ns("train_pytorch_tacotron2.sa.teacher.yaml"))
        else:
            raise KeyError(kind)
    m.eval()
    return m


@contextlib.contextmanager
def prenet_dropout(fn):
    """Replace torch.nn.functional.dropout while the reference decoder runs
    (the only F.dropout call on the inference path is Prenet.forward,
    nets/modules/decoder_sa.py:156-157; nn.Dropout modules call
    torch.nn.functional.dropout too but with training=False, which we pass
    through). `fn(x, p, call_index) -> Tensor`. No reference file is edited."""
    import torch.nn.functional as F
    orig = F.dropout
    state = {"n": 0}

    def patched(input, p=0.5, training=True, inplace=False):
        if not training:
            return orig(input, p, training, inplace)
        i = state["n"]
        state["n"] += 1
        return fn(input, p, i)

    F.dropout = patched
    try:
        yield
    finally:
        F.dropout = orig
