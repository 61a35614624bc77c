"""Hyper-parameters of the FCL-taco2 inference path.

The two presets restate the values of the reference's
conf/train_pytorch_tacotron2.sa.student.yaml:5-18,38-39 (FCL-taco2-S) and
conf/train_pytorch_tacotron2.sa.yaml:5-18,38-39 (FCL-taco2-T); the predictor
and embed sizes are hard-coded in the reference model
(nets/teacher_training/e2e_tts_tacotron2_sa.py:263-286,418-450).
`from_namespace` reads the same argparse.Namespace the reference constructor
receives (model.json / YAML keys with '-' -> '_').
"""
from __future__ import annotations

import argparse
from dataclasses import dataclass, asdict, replace


@dataclass(frozen=True)
class HParams:
    idim: int = 76
    odim: int = 80
    embed_dim: int = 512
    elayers: int = 1
    eunits: int = 512
    econv_layers: int = 3
    econv_chans: int = 512
    econv_filts: int = 5
    dlayers: int = 2
    dunits: int = 1024
    prenet_layers: int = 2
    prenet_units: int = 256
    postnet_layers: int = 5
    postnet_chans: int = 512
    postnet_filts: int = 5
    use_batch_norm: bool = True
    use_concate: bool = True
    use_residual: bool = False
    reduction_factor: int = 1
    dropout_rate: float = 0.5
    zoneout_rate: float = 0.1
    spk_embed_dim: int | None = None
    output_activation: str | None = None
    use_fe_condition: bool = True
    append_position: bool = True
    # predictors: duration from args, pitch/energy hard-coded in the reference
    predictor_layers: int = 2
    predictor_chans: int = 384
    predictor_kernel: int = 3
    embed_kernel: int = 9

    def validate(self) -> None:
        """Reject the combinations SURVEY.md 8(b) lists as out of scope."""
        bad = []
        if self.elayers != 1: bad.append("elayers != 1")
        if self.dlayers != 2: bad.append("dlayers != 2")
        if self.econv_layers != 3: bad.append("econv_layers != 3")
        if self.prenet_layers != 2: bad.append("prenet_layers != 2")
        if self.postnet_layers != 5: bad.append("postnet_layers != 5")
        if self.reduction_factor != 1: bad.append("reduction_factor != 1")
        if self.spk_embed_dim is not None: bad.append("spk_embed_dim")
        if self.use_residual: bad.append("use_residual")
        if self.output_activation is not None: bad.append("output_activation")
        if not self.use_batch_norm: bad.append("use_batch_norm=False")
        if not self.use_concate: bad.append("use_concate=False")
        if not self.use_fe_condition: bad.append("use_fe_condition=False")
        if not self.append_position: bad.append("append_position=False")
        if self.embed_dim != self.econv_chans or self.econv_chans != self.eunits:
            bad.append("embed_dim/econv_chans/eunits must be equal")
        if self.econv_filts != 5 or self.postnet_filts != 5: bad.append("conv filter size != 5")
        if self.predictor_layers != 2 or self.predictor_kernel != 3:
            bad.append("predictor layers/kernel")
        if self.eunits % 64 or self.dunits % 64 or self.postnet_chans % 64 or self.prenet_units % 64:
            bad.append("channel sizes must be multiples of 64")
        if self.zoneout_rate <= 0.0 or self.zoneout_rate >= 1.0: bad.append("zoneout_rate not in (0,1)")
        if bad:
            raise ValueError("unsupported FCL-taco2 configuration for the B200 path: " + ", ".join(bad))

    def to_dict(self):
        return asdict(self)


PRESETS = {
    "T": HParams(),
    "S": HParams(embed_dim=256, eunits=256, econv_chans=256, dunits=256, postnet_chans=128),
}


def preset(name: str, **over) -> HParams:
    return replace(PRESETS[name], **over)


def from_namespace(idim: int, odim: int, args, com_args=None) -> HParams:
    """Build HParams from the Namespace(s) the reference constructor takes
    (nets/teacher_training/e2e_tts_tacotron2_sa.py:289-350)."""
    d = dict(vars(args)) if args is not None else {}
    for k in ("use_fe_condition", "append_position"):
        if k not in d and com_args is not None and hasattr(com_args, k):
            d[k] = getattr(com_args, k)
    base = HParams(idim=idim, odim=odim)
    kw = {}
    for f in base.to_dict():
        if f in ("idim", "odim"):
            continue
        if f in d and d[f] is not None:
            kw[f] = d[f]
    if "duration_predictor_layers" in d: kw["predictor_layers"] = d["duration_predictor_layers"]
    if "duration_predictor_chans" in d: kw["predictor_chans"] = d["duration_predictor_chans"]
    if "duration_predictor_kernel_size" in d: kw["predictor_kernel"] = d["duration_predictor_kernel_size"]
    if d.get("spk_embed_dim") is not None: kw["spk_embed_dim"] = d["spk_embed_dim"]
    if d.get("output_activation") is not None: kw["output_activation"] = d["output_activation"]
    return replace(base, **kw)


def namespace_from_yaml(path: str) -> argparse.Namespace:
    """Read a reference conf/*.yaml into the Namespace form ('-' -> '_')."""
    import yaml
    with open(path) as f:
        y = yaml.safe_load(f)
    return argparse.Namespace(**{k.replace("-", "_"): v for k, v in y.items()})
