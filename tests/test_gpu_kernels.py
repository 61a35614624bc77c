"""Per-kernel parity: each C-ABI entry point vs the oracle restatement on the same seeded inputs.
fp32 path tolerances are stated per test (differences are summation order + BN folding)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from fcl_taco2_b200 import hparams, pack, plan as planmod, synth
from oracle import restate, philox
from tests.helpers import weights, err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engines():
    from fcl_taco2_b200.engine import Engine
    out = {}

    def get(kind, seed=0):
        key = (kind, seed)
        if key not in out:
            hp = hparams.preset(kind)
            sd = weights(kind, seed)
            out[key] = (Engine(hp, pack.pack_fp32(sd, hp), "cuda:0"), sd, hp)
        return out[key]
    return get


def _upload(eng, xs, ds=None):
    pl = planmod.make_plan(xs, ds)
    d, _ = eng.upload(pl)
    return pl, d


def test_len_reg_bit_exact(engines):
    """duration -> frame index map must be bit-exact (north_star); incl. d=0 rows and the stress shape."""
    eng, _, _ = engines("S")
    rs = np.random.RandomState(0)
    for case in range(7):
        if case == 4:                                     # multi-CTA scan + counting sort (P >= 8192): ragged, with zeros
            xs, ds = synth.synth_batch(700, 4)
            for d in ds:
                d[rs.rand(len(d)) < 0.05] = 0
        elif case == 5:                                   # exactly a multiple of the 4096-row block, long durations
            xs, ds = synth.synth_batch(128, 5, fixed_len=128, stress=True)
        elif case == 6:                                   # one row past a block boundary; a duration at the supported maximum
            xs, ds = synth.synth_batch(1, 6, fixed_len=8193)
            ds[0][4096] = 1023
        elif case == 0:
            xs, ds = synth.synth_batch(5, 1)
        elif case == 1:
            xs, ds = synth.synth_batch(3, 2, fixed_len=500, stress=True)
        elif case == 2:                                   # zeros allowed by K0 itself (documented extension)
            xs, ds = synth.synth_batch(4, 3)
            for d in ds:
                d[rs.rand(len(d)) < 0.3] = 0
            ds[0][:] = np.maximum(ds[0], 1)
        else:
            xs, ds = [np.array([5])], [np.array([7])]
        pl, d = _upload(eng, xs, ds)
        frame_off, ufo, order, totals = eng.len_reg_scan(d["dur"], d["utt_off"], pl.n_utts)
        dur = pl.dur.astype(np.int64)
        row, step, off = restate.frame_map(dur)
        F_ = int(off[-1])
        assert totals.cpu().tolist() == [F_, int(dur.max())]
        assert (frame_off.cpu().numpy() == off).all()
        assert (ufo.cpu().numpy() == off[pl.utt_off]).all()
        o = order.cpu().numpy()
        assert sorted(o.tolist()) == list(range(pl.n_rows))
        assert (np.diff(dur[o]) <= 0).all()               # duration-descending
        fmap, pos = eng.frame_map(frame_off, ufo, pl.n_rows, pl.n_utts, F_, want_position=True)
        fm = fmap.cpu().numpy()
        assert (fm[0, :F_] == row).all() and (fm[1, :F_] == step).all()
        utt_of_row = np.repeat(np.arange(pl.n_utts), np.diff(pl.utt_off))
        assert (fm[2, :F_] == off[pl.utt_off][utt_of_row[row]]).all()
        assert (fm[3, :F_] == off[pl.utt_off][utt_of_row[row] + 1]).all()
        ref_pos = (step.astype(np.float32) / dur[row].astype(np.float32))
        assert (pos.cpu().numpy()[:F_] == ref_pos).all()  # IEEE fp32 division, bit-exact


@pytest.mark.parametrize("kind", ["S", "T"])
def test_encoder_and_predictors(engines, kind):
    eng, sd, hp = engines(kind)
    xs, _ = synth.synth_batch(4, 5)
    xs.append(np.array([3]))                              # single-phoneme utterance
    xs.append(np.arange(1, 4))
    pl, d = _upload(eng, xs)
    seg = (d["seg_lo"], d["seg_hi"])
    h = eng.encoder(d["ids"], d["utt_off"], seg, pl.n_utts)
    dlog, dur = eng.predictor("dur", h, seg, want_dur=True)
    pit, _ = eng.predictor("pitch", h, seg)
    ene, _ = eng.predictor("energy", h, seg)
    hn, _ = eng.embed_add(h, pit, ene, seg)
    # packed form: straight into the tensor-core decoder's operand image == fcl_pack_rows_bf16(hn), bit for bit
    from fcl_taco2_b200 import _lib, pack
    from fcl_taco2_b200._lib import dptr
    order = torch.randperm(pl.n_rows, generator=torch.Generator().manual_seed(0)).to(torch.int32).cuda()
    eng.op_dtype = pack.op_dtype()
    hn2, img = eng.embed_add(h, pit, ene, seg, order=order, want_rows=True)
    ref_img = torch.empty_like(img)
    _lib.call("fcl_pack_rows_bf16", _lib.PackRowsParams(n_rows=pl.n_rows, cols=hp.eunits, src=dptr(hn), ld=hp.eunits,
                                                        order=dptr(order), dst=dptr(ref_img)), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert torch.equal(hn2, hn) and torch.equal(img.view(torch.int16), ref_img.view(torch.int16))
    for k, i in enumerate(pl.perm):
        lo, hi = pl.utt_off[k], pl.utt_off[k + 1]
        x = torch.from_numpy(xs[i])
        h_ref = restate.encoder(sd, x)
        assert err(h[lo:hi].cpu(), h_ref)[0] < 2e-5
        dl = restate.predictor_stack(sd, "duration_predictor", h_ref)
        assert err(dlog[lo:hi].cpu(), dl)[0] < 5e-5
        dr = restate.durations_from_log(dl).numpy()
        frac = np.abs((np.exp(dl.numpy()) - 1.0) % 1.0 - 0.5)
        assert ((dur[lo:hi].cpu().numpy() == dr) | (frac < 1e-3)).all()
        p_ref = restate.predictor_stack(sd, "pitch_predictor", h_ref)
        e_ref = restate.predictor_stack(sd, "energy_predictor", h_ref)
        assert err(pit[lo:hi].cpu(), p_ref)[0] < 5e-5 and err(ene[lo:hi].cpu(), e_ref)[0] < 5e-5
        hn_ref = h_ref + restate.embed_scalar(sd, "pitch_embed", p_ref) + restate.embed_scalar(sd, "energy_embed", e_ref)
        assert err(hn[lo:hi].cpu(), hn_ref)[0] < 5e-5


def test_duration_rounding_half_even(engines):
    """clamp(round(exp(x)-1), 0): feed crafted head outputs through the LayerNorm head kernel."""
    eng, _, _ = engines("S")
    C = 384
    vals = torch.tensor([np.log(1.5), np.log(2.5), np.log(3.5), -3.0, np.log(51.0), 8.0], dtype=torch.float32)
    # x rows whose LayerNorm output dotted with head_w gives exactly vals: use gamma=0, beta=1/C*val trick
    x = torch.randn(len(vals), C, device="cuda")
    dur = torch.empty(len(vals), dtype=torch.int32, device="cuda")
    head = torch.empty(len(vals), dtype=torch.float32, device="cuda")
    g = torch.zeros(C, device="cuda")
    w = torch.zeros(C, device="cuda")
    w[0] = 1.0
    for i, v in enumerate(vals.tolist()):
        b = torch.zeros(C, device="cuda")
        b[0] = v
        eng.layernorm(x[i:i + 1], g, b, head_w=w, head_b=0.0, head_out=head[i:i + 1], dur_out=dur[i:i + 1])
    ref = torch.clamp(torch.round(head.cpu().exp() - 1.0), min=0, max=1023).int()
    assert (dur.cpu() == ref).all(), (dur.cpu(), ref)


@pytest.mark.parametrize("kind,tile", [("S", 32), ("S", 16), ("T", 16)])
@pytest.mark.parametrize("drop", [0.0, 0.5])
def test_decoder_steps(engines, kind, tile, drop):
    """K4 vs decoder_sa.py:577-630 restated: same hn, durations, dropout mask -> before_outs."""
    eng, sd, hp = engines(kind)
    xs, ds = synth.synth_batch(3, 9, fixed_len=23)
    pl, d = _upload(eng, xs, ds)
    g = torch.Generator().manual_seed(5)
    hn = torch.randn(pl.n_rows, hp.eunits, generator=g)
    frame_off, ufo, order, totals = eng.len_reg_scan(d["dur"], d["utt_off"], pl.n_utts)
    F_ = int(pl.dur.sum())
    before = eng.decoder(hn.cuda(), d["dur"], frame_off, order, d["row_utt"], d["row_phone"], F_, 0.1, drop, 99,
                         tile_rows=tile)
    torch.cuda.synchronize()
    dur = torch.from_numpy(pl.dur.astype(np.int64))
    steps = restate.decoder_steps(sd, hn, restate.position_table(dur), int(dur.max()), 0.1,
                                  restate.Dropout(drop, 99), pl.row_utt, pl.row_phone)
    row, step, _ = restate.frame_map(dur.numpy())
    ref = steps[torch.from_numpy(row), torch.from_numpy(step)]
    mx, mean = err(before.cpu(), ref)
    assert mx < 1e-4 and mean < 1e-5, (mx, mean)


@pytest.mark.parametrize("kind", ["S", "T"])
def test_postnet(engines, kind):
    eng, sd, hp = engines(kind)
    lens = [37, 5, 1, 64]
    F_ = sum(lens)
    before = torch.randn(F_, 80, generator=torch.Generator().manual_seed(1))
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    lo = torch.from_numpy(np.repeat(off[:-1], lens)).cuda()
    hi = torch.from_numpy(np.repeat(off[1:], lens)).cuda()
    out = eng.postnet(before.cuda(), (lo, hi), F_)
    torch.cuda.synchronize()
    for k in range(len(lens)):
        ref = restate.postnet(sd, before[off[k]:off[k + 1]])
        assert err(out[off[k]:off[k + 1]].cpu(), ref)[0] < 5e-5
