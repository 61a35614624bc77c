"""Parity at BENCHMARK scale: the routes the headline number runs through.

The persistent decoder walks SEVERAL tiles per CTA (pair) through the longest-processing-time schedule, with
TMEM / mbarrier phases carried from one tile to the next and state re-zeroed in between -- none of which the
small-shape tests touch (they give every CTA one tile). Checked here three ways:
  * bit-for-bit against the same utterances decoded in sub-batches that give every CTA exactly ONE tile
    (single-CTA kernel, no group) -- same per-row arithmetic, so `torch.equal`;
  * against the fp32 oracle (oracle/restate.py; decoder_sa.py:577-630) on sampled utterances, stated bf16 bounds;
  * decoder kernel alone: pair route vs single-CTA route vs oracle on sampled ROWS (rows are independent).
Also: T batch 32 through `inference_batch` (group mode g = 7 chosen by the engine) vs the oracle, and the bf16
BiLSTM kernel alone vs `restate.bilstm` in all three tile modes with > 64 tiles.
"""
import numpy as np
import pytest
import torch

from fcl_taco2_b200 import hparams, pack, plan as planmod, synth
from fcl_taco2_b200 import _lib
from fcl_taco2_b200._lib import dptr
from oracle import restate
from tests.helpers import weights, err
from tests.test_gpu_bf16 import DEC_MAX_ABS, DEC_MEAN_L1, MEL_MAX_ABS, MEL_MEAN_L1

pytestmark = pytest.mark.gpu


def _model(kind, seed, drop_seed):
    from fcl_taco2_b200 import model as M
    m = M.from_preset(kind, seed=None, device="cpu", precision="fp16")
    m.load_state_dict(weights(kind, seed))
    return m.to("cuda:0").set_prenet_dropout(rate=0.5, seed=drop_seed)


def test_s_large_batch_default_route_multi_tile():
    """S, 420 utterances (~33 k phoneme rows = ~260 tiles on 148 SMs): engine default = pair kernel, ~3.5 super-tiles
    per CTA pair. Must equal, bit for bit, the same utterances decoded 40 at a time with one tile per CTA."""
    m = _model("S", 0, 1)
    eng = m.engine()
    xs, ds = synth.synth_batch(420, seed=5)
    n_tiles = (sum(len(x) for x in xs) + 127) // 128
    assert n_tiles >= eng.n_slots + 64, "the batch must give every CTA pair several super-tiles"
    assert eng.use_pair is None and eng.force_group == 0            # engine defaults: this is the benchmark's route
    outs = m.inference_batch(xs, durs=ds)
    torch.cuda.synchronize()
    assert all(torch.isfinite(o).all() for o in outs)

    # (a) one tile per CTA, single-CTA kernel, no group: sub-batches of 40 utterances (~25 tiles)
    eng.use_pair, eng.force_group = False, 1
    try:
        for lo in range(0, len(xs), 40):
            sub = m.inference_batch(xs[lo:lo + 40], durs=ds[lo:lo + 40], utt_ids=list(range(lo, min(lo + 40, len(xs)))))
            for i, o in enumerate(sub):
                assert torch.equal(o, outs[lo + i]), (lo + i, float((o - outs[lo + i]).abs().max()))
    finally:
        eng.use_pair, eng.force_group = None, 0

    # (b) the oracle on 16 sampled utterances (first / last of the processing order included: longest and shortest)
    sd = weights("S", 0)
    order = np.argsort([-len(x) for x in xs], kind="stable")
    pick = sorted(set([int(order[0]), int(order[-1])] + np.random.RandomState(0).choice(len(xs), 14, replace=False).tolist()))
    worst = (0.0, 0.0)
    for i in pick:
        ref = restate.inference(sd, torch.from_numpy(xs[i]), dur=ds[i], dropout=restate.Dropout(0.5, 1), utt_index=i,
                                fast_lstm=True)
        mx, mean = err(outs[i].cpu(), ref)
        worst = (max(worst[0], mx), max(worst[1], mean))
    print(f"S batch 420 default route vs oracle ({len(pick)} utterances): max-abs {worst[0]:.3e} mean-L1 {worst[1]:.3e}")
    assert worst[0] < MEL_MAX_ABS and worst[1] < MEL_MEAN_L1, worst


@pytest.mark.parametrize("kind,n_rows", [("S", 148 * 128 * 2 + 300), ("T", 148 * 128 + 64 * 128 + 77)])
def test_decoder_kernel_multi_tile_routes(kind, n_rows):
    """Decoder kernel alone with more tiles than CTAs: pair route (LPT schedule, several super-tiles per pair) ==
    single-CTA route (several tiles per CTA) bit for bit, and both match the oracle on sampled rows."""
    from fcl_taco2_b200.engine import Engine
    hp = hparams.preset(kind)
    sd = weights(kind, 0)
    eng = Engine(hp, pack.pack_fp32(sd, hp), "cuda:0", "fp16")
    rs = np.random.RandomState(3)
    lens = []
    while sum(lens) < n_rows:
        lens.append(int(min(rs.randint(20, 150), n_rows - sum(lens))))
    xs = [synth.phoneme_ids(n, 76, rs) for n in lens]
    ds = [np.clip(1 + rs.poisson(3.0 if kind == "T" else 6.0, size=n), 1, 50 if kind == "S" else 12).astype(np.int64) for n in lens]
    pl = planmod.make_plan(xs, ds)
    d, _ = eng.upload(pl)
    hn = torch.randn(pl.n_rows, hp.eunits, generator=torch.Generator().manual_seed(5))
    hn_d = hn.cuda()
    frame_off, ufo, order, totals = eng.len_reg_scan(d["dur"], d["utt_off"], pl.n_utts)
    F_ = int(pl.dur.sum())
    outs = {}
    for name, pair in (("pair", True), ("single", False)):
        eng.use_pair, eng.force_group = pair, 1
        outs[name] = eng.decoder(hn_d, d["dur"], frame_off, order, d["row_utt"], d["row_phone"], F_, 0.1, 0.5, 99).clone()
    eng.use_pair, eng.force_group = None, 0
    torch.cuda.synchronize()
    assert torch.isfinite(outs["pair"]).all()
    assert torch.equal(outs["pair"], outs["single"]), float((outs["pair"] - outs["single"]).abs().max())
    # oracle on sampled rows: longest and shortest durations (first / last tiles of the sorted order) + random ones
    dur = pl.dur.astype(np.int64)
    by_d = np.argsort(-dur, kind="stable")
    rows = np.unique(np.concatenate([by_d[:48], by_d[-48:], rs.choice(pl.n_rows, 160 if kind == "S" else 64, replace=False)]))
    dsub = torch.from_numpy(dur[rows])
    steps = restate.decoder_steps(sd, hn[rows], restate.position_table(dsub), int(dsub.max()), 0.1,
                                  restate.Dropout(0.5, 99), pl.row_utt[rows], pl.row_phone[rows])
    foff = np.concatenate([[0], np.cumsum(dur)])
    got = outs["pair"].cpu()
    worst = (0.0, 0.0)
    for k, r in enumerate(rows):
        mx, mean = err(got[foff[r]:foff[r + 1]], steps[k, :dur[r]])
        worst = (max(worst[0], mx), max(worst[1], mean))
    print(f"decoder {kind} {pl.n_rows} rows multi-tile vs oracle rows: max-abs {worst[0]:.3e} mean-L1(max over rows) {worst[1]:.3e}")
    assert worst[0] < DEC_MAX_ABS and worst[1] < DEC_MEAN_L1, worst


@pytest.mark.parametrize("kind,max_pairs,n_rows", [("S", 3, 40 * 128 + 50), ("S", 1, 9 * 128), ("S", 5, 23 * 128 + 1), ("T", 2, 13 * 128)])
def test_decoder_two_super_tiles_in_flight(kind, max_pairs, n_rows):
    """A few CTA pairs walking MANY super-tiles with two of them in flight: slots refill at different steps (ragged step
    counts), the list ends with a single active slot, odd tile counts leave the last peer a dummy tile. inflight = 2 must
    equal inflight = 1, the round-1 pair kernel (v1, the engine default) and the single-CTA kernel bit for bit."""
    from fcl_taco2_b200.engine import Engine
    hp = hparams.preset(kind)
    sd = weights(kind, 0)
    eng = Engine(hp, pack.pack_fp32(sd, hp), "cuda:0", "fp16")
    rs = np.random.RandomState(n_rows)
    lens = []
    while sum(lens) < n_rows:
        lens.append(int(min(rs.randint(5, 150), n_rows - sum(lens))))
    xs = [synth.phoneme_ids(n, 76, rs) for n in lens]
    ds = [np.clip(rs.geometric(0.25, size=n), 1, 30 if kind == "S" else 9).astype(np.int64) for n in lens]   # skewed: few long rows
    pl = planmod.make_plan(xs, ds)
    d, _ = eng.upload(pl)
    hn = torch.randn(pl.n_rows, hp.eunits, generator=torch.Generator().manual_seed(5)).cuda()
    frame_off, ufo, order, totals = eng.len_reg_scan(d["dur"], d["utt_off"], pl.n_utts)
    F_ = int(pl.dur.sum())
    outs = {}
    for name, pair, kern, infl, mp in (("single", False, "v1", 1, None), ("v1", True, "v1", 1, max_pairs),
                                       ("v2_1", True, "v2", 1, max_pairs), ("v2_2", True, "v2", 2, max_pairs)):
        eng.use_pair, eng.force_group, eng.pair_kernel, eng.pair_inflight, eng.max_pairs = pair, 1, kern, infl, mp
        outs[name] = eng.decoder(hn, d["dur"], frame_off, order, d["row_utt"], d["row_phone"], F_, 0.1, 0.5, 31).clone()
    eng.use_pair, eng.force_group, eng.pair_kernel, eng.pair_inflight, eng.max_pairs = None, 0, "v1", 2, None
    torch.cuda.synchronize()
    assert torch.isfinite(outs["v2_2"]).all()
    for name in ("v1", "v2_1", "v2_2"):
        assert torch.equal(outs[name], outs["single"]), (name, float((outs[name] - outs["single"]).abs().max()))


def test_t_batch32_group_mode_vs_oracle():
    """BASELINE config 2: FCL-taco2-T, batch 32 (20 tiles: the engine picks group mode, g = 7) vs the oracle."""
    m = _model("T", 0, 7)
    eng = m.engine()
    xs, ds = synth.synth_batch(32, seed=11)
    pl = planmod.make_plan(xs, ds)
    d, _ = eng.upload(pl)
    _, _, order, _ = eng.len_reg_scan(d["dur"], d["utt_off"], pl.n_utts)
    group, n_groups, n_slots, _ = eng.decoder_schedule(order, d["dur"], pl.n_rows)
    assert group >= 4, f"expected group mode for {(pl.n_rows + 127) // 128} tiles, got group {group}"
    outs = m.inference_batch(xs, durs=ds)
    torch.cuda.synchronize()
    sd = weights("T", 0)
    worst = (0.0, 0.0)
    for i in (0, 9, 17, 31):
        ref = restate.inference(sd, torch.from_numpy(xs[i]), dur=ds[i], dropout=restate.Dropout(0.5, 7), utt_index=i,
                                fast_lstm=True)
        mx, mean = err(outs[i].cpu(), ref)
        worst = (max(worst[0], mx), max(worst[1], mean))
    print(f"T batch 32 (group {group}) vs oracle: max-abs {worst[0]:.3e} mean-L1 {worst[1]:.3e}")
    assert worst[0] < MEL_MAX_ABS and worst[1] < MEL_MEAN_L1, worst
    # and the group-mode batch equals a single-utterance call bit for bit
    single = m.inference(torch.from_numpy(xs[9]), None, dur=ds[9], dropout_utt_index=9)
    assert torch.equal(single, outs[9])


@pytest.mark.parametrize("kind,n_utts,tile_utts", [("S", 2200, 32), ("S", 4300, 64), ("S", 8400, 128), ("T", 300, 32)])
def test_bilstm_bf16_kernel_many_tiles(kind, n_utts, tile_utts):
    """fcl_bilstm_bf16 alone, > 64 tiles in every tile mode, vs restate.bilstm on the SAME bf16-rounded input
    projection (what the kernel reads): isolates the recurrence (bf16 h operand, tanh.approx gates)."""
    from fcl_taco2_b200.engine import Engine
    hp = hparams.preset(kind)
    sd = weights(kind, 0)
    packed = pack.pack_fp32(sd, hp)
    E, hd = hp.eunits, hp.eunits // 2
    rs = np.random.RandomState(n_utts)
    lens = np.sort(rs.randint(1, 24, size=n_utts))[::-1].copy()        # longest first, as the planner orders them
    lens[0] = 40
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    P = int(off[-1])
    x = torch.randn(P, hp.econv_chans, generator=torch.Generator().manual_seed(1))
    # gate-interleaved input projection, rounded to bf16 (the kernel's gx operand)
    gx = (x @ packed["blstm_wih"][0] + packed["blstm_b"]).to(pack.op_dtype())
    whh = pack.pack_bilstm_whh_bf16(packed).cuda()
    n_tiles = (n_utts + tile_utts - 1) // tile_utts
    assert n_tiles > 64 or kind == "T"
    c_ws = torch.empty(n_tiles * 2 * hd * 128, dtype=torch.float32, device="cuda")
    out = torch.full((P, E), float("nan"), device="cuda")
    gx_d, off_d = gx.cuda(), torch.from_numpy(off).cuda()
    _lib.call("fcl_bilstm_bf16", _lib.BiLstmBf16Params(n_utts=n_utts, hidden=hd, tile_utts=tile_utts, utt_off=dptr(off_d),
                                                        gx=dptr(gx_d), whh_packed=dptr(whh), c_ws=dptr(c_ws), out=dptr(out)),
              torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = out.cpu()
    assert torch.isfinite(got).all()
    # oracle recurrence on the bf16-rounded projection: de-interleave gx back to torch's [i;f;g;o] row blocks
    gxf = gx.float().view(P, 2, hd, 4).permute(0, 1, 3, 2).reshape(P, 2, 4 * hd)
    pick = np.unique(np.concatenate([[0, 1, n_utts - 1, n_utts - 2], rs.choice(n_utts, 40, replace=False)]))
    worst = 0.0
    for u in pick:
        n = int(lens[u])
        for di, suf in enumerate(("", "_reverse")):
            w = sd["enc.blstm.weight_hh_l0" + suf]
            h = torch.zeros(hd); c = torch.zeros(hd)
            ts = range(n) if di == 0 else range(n - 1, -1, -1)
            for t in ts:
                g = gxf[off[u] + t, di] + w @ h
                i, f, gg, o = g[:hd], g[hd:2 * hd], g[2 * hd:3 * hd], g[3 * hd:]
                c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
                h = torch.sigmoid(o) * torch.tanh(c)
                worst = max(worst, float((got[off[u] + t, di * hd:(di + 1) * hd] - h).abs().max()))
    print(f"bilstm bf16 {kind} {n_utts} utts tile {tile_utts}: max-abs vs oracle recurrence {worst:.3e}")
    assert worst < 1.5e-2, worst
