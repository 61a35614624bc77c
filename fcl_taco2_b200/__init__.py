"""fcl_taco2_b200: B200-native (sm_100a) inference path of FCL-taco2.

Only the hot path named in SURVEY.md section 8: `Tacotron2_sa.inference()` and its
batched form, hand-written CUDA behind the C ABI of include/fcl_taco2.h.
There is no CPU fallback: importing `fcl_taco2_b200.model` without the built
shared library raises.
"""
from .hparams import HParams, PRESETS, preset  # noqa: F401

__version__ = "0.1.0"
