"""oracle/restate.py against the imported reference modules (build container only: needs
/root/reference; skipped on the GPU box)."""
import numpy as np
import pytest
import torch

from fcl_taco2_b200 import hparams, synth
from oracle import ref_loader, restate
from tests.helpers import weights, err

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")


@pytest.mark.parametrize("kind", ["S", "T"])
def test_restatement_vs_reference_modules(kind):
    torch.manual_seed(0)
    m = ref_loader.build(kind)
    sd = weights(kind, 3)
    m.load_state_dict(sd, strict=True)
    xs, ds = synth.synth_batch(1, 11, fixed_len=24)
    x, d = torch.from_numpy(xs[0]), torch.from_numpy(ds[0])
    m.dec.prenet.dropout_rate = 0.0
    with torch.no_grad():
        ref = m.inference(x, None, dur=d)
        # forced f0 / energy path (e2e_tts_tacotron2_sa.py:649-651)
        f0 = torch.randn(24, 1)
        en = torch.randn(24, 1)
        ref_forced = m.inference(x, None, dur=d, f0=f0, energy=en)
    mine = restate.inference(sd, x, dur=d)
    assert err(ref, mine)[0] < 5e-5
    mine_forced = restate.inference(sd, x, dur=d, f0=f0, energy=en)
    assert err(ref_forced, mine_forced)[0] < 5e-5
    assert err(mine, restate.inference(sd, x, dur=d, fast_lstm=True))[0] < 5e-5


def test_reference_default_dropout_is_stochastic():
    """Documents SURVEY.md fact 3: the reference's own output is random at inference."""
    m = ref_loader.build("S")
    m.load_state_dict(weights("S", 0), strict=True)
    xs, ds = synth.synth_batch(1, 0, fixed_len=16)
    x, d = torch.from_numpy(xs[0]), torch.from_numpy(ds[0])
    with torch.no_grad():
        a = m.inference(x, None, dur=d)
        b = m.inference(x, None, dur=d)
    assert err(a, b)[0] > 1e-3
