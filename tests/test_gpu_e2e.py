"""End-to-end parity through the public model API (C ABI underneath).

fp32 path tolerance vs the reference's own outputs (tests/golden) and vs the oracle:
max-abs <= 2e-4, mean-L1 <= 2e-5 on mels of abs-mean ~1 (observed values are printed by
bench/smoke; the bound leaves room for summation-order differences over 40 AR steps)."""
import numpy as np
import pytest
import torch

from fcl_taco2_b200 import synth
from oracle import restate
from tests.conftest import golden_cases
from tests.helpers import load_golden, weights, err

pytestmark = pytest.mark.gpu
MAX_ABS, MEAN_L1 = 2e-4, 2e-5


@pytest.fixture(scope="module")
def models():
    from fcl_taco2_b200 import model as M
    cache = {}

    def get(kind, seed):
        if (kind, seed) not in cache:
            m = M.from_preset(kind, seed=None, device="cpu", kd_keys=(kind == "S"))
            m.load_state_dict(weights(kind, seed), strict=True)
            cache[(kind, seed)] = m.to("cuda:0")
        return cache[(kind, seed)]
    return get


@pytest.mark.parametrize("name", golden_cases())
def test_against_reference_golden(models, name):
    g = load_golden(name)
    m = models(g["kind"], g["weight_seed"])
    assert synth.state_dict_digest(weights(g["kind"], g["weight_seed"])) == g["weights_sha256"]
    m.set_prenet_dropout(rate=g["dropout_rate"], seed=g["dropout_seed"])
    out = m.inference(torch.from_numpy(g["x"]).cuda(), None, dur=torch.from_numpy(g["dur"]).cuda(),
                      dropout_utt_index=g["utt_index"])
    assert out.is_cuda and out.dtype == torch.float32 and tuple(out.shape) == g["out"].shape
    mx, mean = err(out.cpu(), g["out"])
    assert mx < MAX_ABS and mean < MEAN_L1, (mx, mean)


@pytest.mark.parametrize("kind", ["S", "T"])
def test_batched_equals_looped_and_oracle(models, kind):
    m = models(kind, 1)
    m.set_prenet_dropout(rate=0.5, seed=4242)
    xs, ds = synth.synth_batch(6 if kind == "S" else 3, 21)
    outs = m.inference_batch(xs, durs=ds)
    sd = weights(kind, 1)
    for i in range(len(xs)):
        single = m.inference(torch.from_numpy(xs[i]), None, dur=ds[i], dropout_utt_index=i)
        assert torch.equal(single, outs[i]), "batched result must equal the single-utterance call bit for bit"
        ref = restate.inference(sd, torch.from_numpy(xs[i]), dur=ds[i], dropout=restate.Dropout(0.5, 4242), utt_index=i)
        mx, mean = err(outs[i].cpu(), ref)
        assert mx < MAX_ABS and mean < MEAN_L1, (i, mx, mean)


def test_forced_f0_energy_and_predicted_durations(models):
    m = models("S", 2)
    m.set_prenet_dropout(rate=0.0)
    sd = weights("S", 2)
    xs, ds = synth.synth_batch(2, 33, fixed_len=30)
    rs = np.random.RandomState(0)
    f0 = [rs.randn(30).astype(np.float32) for _ in xs]
    en = [rs.randn(30).astype(np.float32) for _ in xs]
    outs = m.inference_batch(xs, durs=ds, f0s=f0, energies=en)
    for i in range(2):
        ref = restate.inference(sd, torch.from_numpy(xs[i]), dur=ds[i], f0=f0[i], energy=en[i])
        assert err(outs[i].cpu(), ref)[0] < MAX_ABS
    # predicted durations: random-init predictors emit zeros -> same error class as the reference's assert
    with pytest.raises(ValueError):
        m.inference_batch(xs)


def test_predicted_duration_path(models):
    from fcl_taco2_b200 import model as M
    sd = dict(weights("S", 2))
    sd["duration_predictor.linear.bias"] = torch.tensor([1.6])
    m = M.from_preset("S", seed=None, device="cpu")
    m.load_state_dict(sd)
    m = m.to("cuda:0").set_prenet_dropout(rate=0.0)
    xs, _ = synth.synth_batch(3, 44, fixed_len=20)
    outs = m.inference_batch(xs)
    for i in range(3):
        r = restate.inference(sd, torch.from_numpy(xs[i]), return_all=True)
        assert outs[i].shape == r["out"].shape        # same predicted durations
        assert err(outs[i].cpu(), r["out"])[0] < MAX_ABS


def test_stress_500_phonemes_batch(models):
    """config 4: 500 phonemes, skewed durations up to 40, ragged masking; batch of 3 + short ones."""
    m = models("S", 2)
    m.set_prenet_dropout(rate=0.0)
    xs, ds = synth.synth_batch(2, 55, fixed_len=500, stress=True)
    xs2, ds2 = synth.synth_batch(3, 56)
    xs, ds = xs + xs2, ds + ds2
    outs = m.inference_batch(xs, durs=ds)
    sd = weights("S", 2)
    for i in (0, 3):
        ref = restate.inference(sd, torch.from_numpy(xs[i]), dur=ds[i])
        mx, mean = err(outs[i].cpu(), ref)
        assert mx < MAX_ABS and mean < MEAN_L1, (mx, mean)
    for i in range(len(xs)):
        assert outs[i].shape == (int(ds[i].sum()), 80) and torch.isfinite(outs[i]).all()


def test_errors_are_loud(models):
    m = models("S", 2)
    with pytest.raises(ValueError):
        m.inference(torch.tensor([1, 2, 3]), None, dur=torch.tensor([1, 0, 2]))      # zero duration
    with pytest.raises(ValueError):
        m.inference(torch.tensor([1, 2, 300]), None, dur=torch.tensor([1, 1, 2]))    # id out of range
    with pytest.raises(ValueError):
        m.inference_batch([np.array([1, 2])], durs=[np.array([1])])                  # ragged mismatch
    with pytest.raises(ValueError):
        m.inference(torch.tensor([1, 2]), None, spemb=torch.zeros(4), dur=torch.tensor([1, 1]))
    with pytest.raises(ValueError):                                                   # forward() without durations / prosody
        m.forward(torch.tensor([[1, 2]]), torch.tensor([2]), torch.zeros(1, 3, 80), torch.tensor([3]))


@pytest.mark.gpu
def test_inference_stream_matches_inference_batch():
    """The pipelined API (D2H of batch i overlapped with batch i+1) returns exactly what inference_batch returns,
    batch by batch, also when the batches differ in size (the pinned staging ring is reused and regrown)."""
    from fcl_taco2_b200 import model as M, synth
    m = M.from_preset("S", seed=0, device="cuda:0", precision="fp16").set_prenet_dropout(rate=0.5, seed=3)
    sets = [synth.synth_batch(n, seed) for n, seed in ((5, 1), (9, 2), (3, 3), (12, 4), (1, 5))]
    want = []
    for k, (xs, ds) in enumerate(sets):
        outs = m.inference_batch(xs, durs=ds, utt_ids=list(range(100 * k, 100 * k + len(xs))))
        want.append([o.cpu().numpy().copy() for o in outs])
    batches = ({"xs": xs, "durs": ds, "utt_ids": list(range(100 * k, 100 * k + len(xs)))} for k, (xs, ds) in enumerate(sets))
    n = 0
    for k, outs in enumerate(m.inference_stream(batches)):
        assert len(outs) == len(want[k])
        for a, b in zip(outs, want[k]):
            assert a.shape == b.shape and np.array_equal(a, b)
        n += 1
    assert n == len(sets)


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_zero_duration_phonemes_are_skipped(precision):
    """Extension: d = 0 phonemes feed the encoder / predictors / prosody embeddings but produce no
    decoder row and no frames, as in the reference's forward() (decoder_sa.py:459-463). Without the opt-in the call
    raises like the reference's assert (decoder_sa.py:575). Also: durations above FCL_MAX_DURATION raise (no clamp)."""
    from fcl_taco2_b200 import model as M
    m = M.from_preset("S", seed=None, device="cpu", precision=precision)
    m.load_state_dict(weights("S", 4))
    m = m.to("cuda:0").set_prenet_dropout(rate=0.5, seed=21)
    xs, ds = synth.synth_batch(5, 91, fixed_len=37)
    ds = [d.copy() for d in ds]
    ds[0][[0, 5, 6, 7, 36]] = 0                 # first, a run, last
    ds[1][::2] = 0                              # every other phoneme
    ds[2][:] = 0; ds[2][18] = 4                 # a single voiced phoneme
    ds[3][:] = 0                                # an utterance that produces no frames at all
    with pytest.raises(ValueError):
        m.inference_batch(xs, durs=ds)
    outs = m.inference_batch(xs, durs=ds, skip_zero_durations=True)
    sd = weights("S", 4)
    tol = (MAX_ABS, MEAN_L1) if precision == "fp32" else (3e-2, 4e-3)
    for i in range(5):
        assert outs[i].shape == (int(ds[i].sum()), 80)
        if ds[i].sum() == 0:
            continue
        ref = restate.inference(sd, torch.from_numpy(xs[i]), dur=ds[i], dropout=restate.Dropout(0.5, 21), utt_index=i,
                                skip_zero=True)
        mx, mean = err(outs[i].cpu(), ref)
        assert mx < tol[0] and mean < tol[1], (i, mx, mean)
    big = [d.copy() for d in ds]
    big[4][3] = 1024
    with pytest.raises(ValueError, match="exceeds the supported maximum"):
        m.inference_batch(xs, durs=big, skip_zero_durations=True)
