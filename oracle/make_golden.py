"""Generate tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE (build container only).

    python -m oracle.make_golden

TEST INFRASTRUCTURE. The reference has no golden vectors of its own
(SURVEY.md section 4), and it is Python, so it cannot travel to the GPU box: these
fixtures are what pins both oracle/restate.py and the CUDA path to it there.
Each case: seeded weights (fcl_taco2_b200.synth.random_state_dict -- regenerated
on the other side, sha256 stored), ids, forced durations, and the outputs of
the reference's Encoder.inference / predictors / Tacotron2_sa.inference.
Prenet dropout is either disabled (rate 0) or replaced by the counter-based
mask of oracle/philox.py via oracle.ref_loader.prenet_dropout.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader, philox            # noqa: E402
from fcl_taco2_b200 import synth, hparams        # noqa: E402

CASES = [
    # name, kind, weight seed, input seed, N, stress, dropout rate, dropout seed, utt index
    ("S_n80_nodrop", "S", 0, 0, 80, False, 0.0, 0, 0),
    ("S_n80_drop", "S", 1, 1, 80, False, 0.5, 20211, 5),
    ("S_n500_stress", "S", 2, 2, 500, True, 0.0, 0, 0),
    ("T_n40_nodrop", "T", 0, 3, 40, False, 0.0, 0, 0),
    ("T_n40_drop", "T", 1, 4, 40, False, 0.5, 77, 2),
    ("S_n1_single", "S", 0, 5, 1, False, 0.0, 0, 0),
]


def run_case(kind, wseed, iseed, n, stress, rate, dseed, utt):
    hp = hparams.preset(kind)
    sd = synth.random_state_dict(hp, wseed, kind == "S", hparams.preset("T"))
    m = ref_loader.build(kind)
    m.load_state_dict(sd, strict=True)
    xs, ds = synth.synth_batch(1, iseed, fixed_len=n, stress=stress)
    x, d = torch.from_numpy(xs[0]), torch.from_numpy(ds[0])
    with torch.no_grad():
        h = m.enc.inference(x)
        dlog = m.duration_predictor(h.unsqueeze(0), None)[0]
        dpred = m.duration_predictor.inference(h.unsqueeze(0), None)[0]
        p_out = m.pitch_predictor(h.unsqueeze(0), None)[0, :, 0]
        e_out = m.energy_predictor(h.unsqueeze(0), None)[0, :, 0]
        m.dec.prenet.dropout_rate = rate
        if rate == 0.0:
            out = m.inference(x, None, dur=d)
        else:
            nn = x.shape[0]

            def fn(inp, p, i):
                keep = philox.keep_mask(dseed, np.full(nn, utt), np.arange(nn), i // 2, i % 2, inp.shape[1], p)
                return inp * (torch.from_numpy(keep).float() * (1.0 / (1.0 - p)))
            with ref_loader.prenet_dropout(fn):
                out = m.inference(x, None, dur=d)
    return dict(kind=kind, weight_seed=wseed, weights_sha256=synth.state_dict_digest(sd),
                x=x.numpy(), dur=d.numpy(), h=h.numpy(), dlog=dlog.numpy(), dpred=dpred.numpy(),
                p_out=p_out.numpy(), e_out=e_out.numpy(), out=out.numpy(),
                dropout_rate=rate, dropout_seed=dseed, utt_index=utt)


def main():
    assert ref_loader.available(), "needs /root/reference"
    outdir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    os.makedirs(outdir, exist_ok=True)
    torch.manual_seed(0)
    for name, *cfg in CASES:
        r = run_case(*cfg)
        np.savez_compressed(os.path.join(outdir, name + ".npz"), **r)
        print(name, r["out"].shape, float(np.abs(r["out"]).mean()))


if __name__ == "__main__":
    main()
