"""oracle/restate.py against the outputs of the reference itself (tests/golden, made by
oracle/make_golden.py). Runs anywhere -- no /root/reference needed. fp32 tolerance 5e-5
(observed 6e-6: differences are summation order only)."""
import numpy as np
import pytest
import torch

from fcl_taco2_b200 import synth
from oracle import restate
from tests.conftest import golden_cases
from tests.helpers import load_golden, weights, err

TOL = 5e-5


@pytest.mark.parametrize("name", golden_cases())
def test_restatement_matches_reference_golden(name):
    g = load_golden(name)
    sd = weights(g["kind"], g["weight_seed"])
    assert synth.state_dict_digest(sd) == g["weights_sha256"], "seeded weights drifted (torch RNG changed?)"
    r = restate.inference(sd, torch.from_numpy(g["x"]), dur=g["dur"], return_all=True,
                          dropout=restate.Dropout(g["dropout_rate"], g["dropout_seed"]), utt_index=g["utt_index"])
    assert r["out"].shape == g["out"].shape == (int(g["dur"].sum()), 80)
    assert err(r["h"], g["h"])[0] < TOL
    assert err(r["dlog"], g["dlog"])[0] < TOL
    assert err(r["p_out"], g["p_out"])[0] < TOL
    assert err(r["e_out"], g["e_out"])[0] < TOL
    mx, mean = err(r["out"], g["out"])
    assert mx < TOL and mean < 5e-6, (mx, mean)
    # integer durations: identical unless exp(dlog)-1 sits on a rounding boundary
    dp = restate.durations_from_log(r["dlog"]).numpy()
    frac = np.abs((np.exp(g["dlog"]) - 1.0) % 1.0 - 0.5)
    assert ((dp == g["dpred"]) | (frac < 1e-4)).all()


def test_position_table_and_frame_map_exact():
    d = torch.tensor([3, 1, 5, 2])
    pos = restate.position_table(d)
    assert pos.shape == (4, 5)
    assert pos[0].tolist() == [0.0, np.float32(1) / np.float32(3), np.float32(2) / np.float32(3), 0.0, 0.0]
    assert pos[2, 4] == np.float32(4) / np.float32(5)
    row, step, off = restate.frame_map(d.numpy())
    assert row.tolist() == [0, 0, 0, 1, 2, 2, 2, 2, 2, 3, 3]
    assert step.tolist() == [0, 1, 2, 0, 0, 1, 2, 3, 4, 0, 1]
    assert off.tolist() == [0, 3, 4, 9, 11]


def test_zero_duration_rejected():
    g = load_golden("S_n1_single")
    sd = weights("S", 0)
    with pytest.raises(ValueError):
        restate.inference(sd, torch.from_numpy(g["x"]), dur=np.zeros(1, dtype=np.int64))
