"""Tensor-core (tcgen05) path parity.

Kernel-level: fcl_conv_gemm_bf16 vs a torch fp32 GEMM on the SAME bf16-rounded operands -- isolates
layout/descriptor bugs from precision (tolerance 2e-3 relative to the output scale: fp32 accumulation
order only). Model-level bf16 tolerances are in test_gpu_e2e_bf16.py."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from fcl_taco2_b200 import _lib, pack
from fcl_taco2_b200._lib import dptr

pytestmark = pytest.mark.gpu


def _ref_conv(a, w, bias, seg_lens, taps, act, residual):
    """a (rows,cin) fp32, w (taps,cin,cout) -> per-segment conv with zero halo, operands rounded to bf16."""
    ab = a.to(pack.op_dtype()).float()
    wb = w.to(pack.op_dtype()).float()
    outs, o = [], 0
    for n in seg_lens:
        x = ab[o:o + n].t().unsqueeze(0)                                   # (1,cin,n)
        y = F.conv1d(x, wb.permute(2, 1, 0).contiguous(), bias, 1, taps // 2)[0].t()
        outs.append(y)
        o += n
    y = torch.cat(outs)
    if act == 1:
        y = torch.relu(y)
    elif act == 2:
        y = torch.tanh(y)
    if residual is not None:
        y = y + residual
    return y


@pytest.mark.parametrize("rows,cin,cout,taps,act,use_res,use_gather", [
    (300, 256, 256, 5, 1, False, False),
    (1000, 80, 128, 5, 2, False, False),      # kstage 80
    (257, 384, 384, 3, 1, False, False),      # ntile 192
    (128, 256, 1024, 1, 0, False, False),     # 4 column tiles
    (77, 128, 80, 5, 0, True, False),         # ntile 80 + residual
    (513, 512, 512, 5, 1, False, True),       # embedding gather, 2 column tiles
    (5, 64, 16, 1, 0, False, False),          # tiny
])
def test_conv_gemm_bf16_matches_bf16_rounded_reference(rows, cin, cout, taps, act, use_res, use_gather):
    _lib.load()
    g = torch.Generator().manual_seed(rows + cin)
    w = torch.randn(taps, cin, cout, generator=g) / np.sqrt(cin * taps)
    bias = torch.randn(cout, generator=g) * 0.1
    seg_lens, left = [], rows
    while left > 0:
        n = min(left, int(torch.randint(1, 90, (1,), generator=g)))
        seg_lens.append(n)
        left -= n
    off = np.concatenate([[0], np.cumsum(seg_lens)]).astype(np.int32)
    lo = torch.from_numpy(np.repeat(off[:-1], seg_lens)).cuda()
    hi = torch.from_numpy(np.repeat(off[1:], seg_lens)).cuda()
    if use_gather:
        table = torch.randn(76, cin, generator=g)
        ids = torch.randint(0, 76, (rows,), generator=g)
        a_dev, a_ref, gather = table.cuda(), table[ids], ids.cuda()
    else:
        a_ref = torch.randn(rows, cin, generator=g)
        a_dev, gather = a_ref.cuda(), None
    res = torch.randn(rows, cout, generator=g) if use_res else None
    wp, ntile, kstage = pack.pack_conv_bf16(w)
    wp = wp.cuda()
    out = torch.full((rows, cout), float("nan"), device="cuda")
    bias_d = bias.cuda()
    res_d = res.cuda() if use_res else None
    stream = torch.cuda.current_stream().cuda_stream
    tsrc = tdst = cnt = None
    max_tiles = 0
    if taps > 1:                                      # tile maps: a tile never crosses a segment
        seg_off = torch.from_numpy(off).cuda()
        max_tiles = int(sum((n + 127) // 128 for n in seg_lens)) + 3        # + slack: extra CTAs must exit cleanly
        first = torch.empty(len(seg_lens) + 1, dtype=torch.int32, device="cuda")
        tsrc = torch.empty((max_tiles, 136), dtype=torch.int32, device="cuda")
        tdst = torch.empty((max_tiles, 128), dtype=torch.int32, device="cuda")
        cnt = torch.empty(1, dtype=torch.int32, device="cuda")
        _lib.call("fcl_conv_tiles", _lib.ConvTilesParams(n_segs=len(seg_lens), max_tiles=max_tiles, halo=2,
                                                         seg_off=dptr(seg_off), seg_first_tile=dptr(first),
                                                         tile_src=dptr(tsrc), tile_dst=dptr(tdst), n_tiles=dptr(cnt)), stream)
    p = _lib.ConvGemmBf16Params(rows=rows, cin=cin, cout=cout, taps=taps, a=dptr(a_dev), lda=cin, gather=dptr(gather),
                                tile_src=dptr(tsrc), tile_dst=dptr(tdst), n_tiles_dev=dptr(cnt), n_tiles=max_tiles,
                                map_halo=2, w_packed=dptr(wp), ntile=ntile, kstage=kstage, bias=dptr(bias_d),
                                residual=dptr(res_d), ldr=cout, out=dptr(out), ldo=cout, act=act)
    _lib.call("fcl_conv_gemm_bf16", p, stream)
    torch.cuda.synchronize()
    ref = _ref_conv(a_ref, w, bias, seg_lens if taps > 1 else [rows], taps, act, res)
    got = out.cpu()
    assert torch.isfinite(got).all()
    err = float((got - ref).abs().max())
    assert err < 3e-3 * max(1.0, float(ref.abs().max())), err     # tanh.approx in the epilogue: ~1e-3


# ----------------------------------------------------------------------------- decoder / end-to-end, bf16 path
# Stated tolerances of the tensor-core path (bf16 GEMM operands, fp32 accumulate, fp32 cell state,
# tanh.approx activations) against the fp32 oracle, on mels of abs-mean ~1:
DEC_MAX_ABS, DEC_MEAN_L1 = 1e-2, 1.5e-3     # decoder output before the postnet
MEL_MAX_ABS, MEL_MEAN_L1 = 3e-2, 4e-3       # final mels, every GEMM on the tensor cores with fp16 operands (observed
                                            # <= 1.3e-2 / 1.2e-3 on S-1024, T-32 and the 500-phoneme stress batch; with
                                            # bf16 operands -- round 1 -- the same path measured 8e-2 / 8.7e-3)

from fcl_taco2_b200 import hparams, plan as planmod, synth          # noqa: E402
from oracle import restate                                         # noqa: E402
from tests.conftest import golden_cases                            # noqa: E402
from tests.helpers import weights, err, load_golden                # noqa: E402


@pytest.fixture(scope="module")
def bf16_engines():
    from fcl_taco2_b200.engine import Engine
    out = {}

    def get(kind, seed=0):
        if (kind, seed) not in out:
            hp = hparams.preset(kind)
            sd = weights(kind, seed)
            out[(kind, seed)] = (Engine(hp, pack.pack_fp32(sd, hp), "cuda:0", "fp16"), sd, hp)
        return out[(kind, seed)]
    return get


@pytest.mark.parametrize("kind,n_utts", [("S", 3), ("S", 9), ("T", 2)])
@pytest.mark.parametrize("drop", [0.0, 0.5])
def test_decoder_bf16_steps(bf16_engines, kind, n_utts, drop):
    eng, sd, hp = bf16_engines(kind)
    xs, ds = synth.synth_batch(n_utts, 9, fixed_len=47)          # 141 / 423 / 94 rows: partial tiles, >1 tile
    pl = planmod.make_plan(xs, ds)
    d, _ = eng.upload(pl)
    hn = torch.randn(pl.n_rows, hp.eunits, generator=torch.Generator().manual_seed(5))
    frame_off, ufo, order, totals = eng.len_reg_scan(d["dur"], d["utt_off"], pl.n_utts)
    F_ = int(pl.dur.sum())
    before = eng.decoder(hn.cuda(), d["dur"], frame_off, order, d["row_utt"], d["row_phone"], F_, 0.1, drop, 99)
    torch.cuda.synchronize()
    dur = torch.from_numpy(pl.dur.astype(np.int64))
    steps = restate.decoder_steps(sd, hn, restate.position_table(dur), int(dur.max()), 0.1,
                                  restate.Dropout(drop, 99), pl.row_utt, pl.row_phone)
    row, step, _ = restate.frame_map(dur.numpy())
    ref = steps[torch.from_numpy(row), torch.from_numpy(step)]
    assert torch.isfinite(before).all()
    mx, mean = err(before.cpu(), ref)
    print(f"decoder bf16 {kind} drop={drop}: max-abs {mx:.3e} mean-L1 {mean:.3e}")
    assert mx < DEC_MAX_ABS and mean < DEC_MEAN_L1, (mx, mean)


@pytest.mark.parametrize("name", golden_cases())
def test_e2e_bf16_against_reference_golden(name):
    from fcl_taco2_b200 import model as M
    g = load_golden(name)
    m = M.from_preset(g["kind"], seed=None, device="cpu", precision="fp16")
    m.load_state_dict(weights(g["kind"], g["weight_seed"]))
    m = m.to("cuda:0").set_prenet_dropout(rate=g["dropout_rate"], seed=g["dropout_seed"])
    out = m.inference(torch.from_numpy(g["x"]), None, dur=g["dur"], dropout_utt_index=g["utt_index"])
    mx, mean = err(out.cpu(), g["out"])
    print(f"e2e bf16 {name}: max-abs {mx:.3e} mean-L1 {mean:.3e}")
    assert mx < MEL_MAX_ABS and mean < MEL_MEAN_L1, (mx, mean)


def test_bf16_batched_equals_looped():
    from fcl_taco2_b200 import model as M
    m = M.from_preset("S", seed=3, device="cuda:0", precision="fp16").set_prenet_dropout(rate=0.5, seed=8)
    xs, ds = synth.synth_batch(5, 77)
    outs = m.inference_batch(xs, durs=ds)
    for i in range(len(xs)):
        single = m.inference(torch.from_numpy(xs[i]), None, dur=ds[i], dropout_utt_index=i)
        assert torch.equal(single, outs[i])


def test_fused_postnet_stack_matches_layer_by_layer(bf16_engines):
    """fcl_conv_stack_bf16 (activations stay on chip) vs five fcl_conv_gemm_bf16 launches: same bf16 roundings,
    so they agree to fp32 summation order; utterance boundaries inside and across tiles, 1-frame utterance."""
    eng, sd, hp = bf16_engines("S")
    lens = [300, 5, 1, 112, 113, 64, 700]
    F_ = sum(lens)
    before = torch.randn(F_, 80, generator=torch.Generator().manual_seed(1)).cuda()
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    lo = torch.from_numpy(np.repeat(off[:-1], lens)).cuda()
    hi = torch.from_numpy(np.repeat(off[1:], lens)).cuda()
    ufo = torch.from_numpy(off).cuda()
    tiles = eng.conv_tiles(ufo, len(lens), sum((n + 127) // 128 for n in lens))
    ref = eng.postnet(before, (lo, hi, tiles), F_)                          # layer by layer
    eng.use_img_postnet = False
    fused = eng.postnet(before, (lo, hi, tiles, (ufo, len(lens))), F_)      # fused stack
    eng.use_img_postnet = True
    img = eng.postnet(before, (lo, hi, tiles, (ufo, len(lens))), F_)        # five image-to-image launches (engine default)
    torch.cuda.synchronize()
    assert torch.isfinite(img).all()
    mx_i, mean_i = err(img.cpu(), ref.cpu())
    assert mx_i < 6e-3 and mean_i < 4e-4, (mx_i, mean_i)                    # same bf16 roundings, different summation order
    for k in range(len(lens)):
        o = restate.postnet(sd, before[off[k]:off[k + 1]].cpu())
        assert err(img[off[k]:off[k + 1]].cpu(), o)[0] < 2e-2
    assert torch.isfinite(fused).all()
    # The fused stack pads the first layer's K to 128 (different fp32 summation order): a handful of layer-0 outputs
    # round to the neighbouring bf16 value and each flip fans out over +-8 rows by the last layer (tools/dbg_stack.py:
    # both paths sit at the same distance from the fp32 oracle). Everything else is bit-identical.
    mx, mean = err(fused.cpu(), ref.cpu())
    assert mx < 6e-3 and mean < 4e-4, (mx, mean)
    for k in range(len(lens)):
        o = restate.postnet(sd, before[off[k]:off[k + 1]].cpu())
        assert err(fused[off[k]:off[k + 1]].cpu(), o)[0] < 2e-2


def test_chunked_postnet_bit_identical(bf16_engines):
    """Postnet launched per group of utterances (so that a gather / D2H of finished frames can overlap the rest):
    tiles never cross utterances, so the result must equal the single-launch stack bit for bit."""
    eng, sd, hp = bf16_engines("S")
    lens = [300, 5, 1, 112, 113, 64, 700, 9, 250]
    F_ = sum(lens)
    before = torch.randn(F_, 80, generator=torch.Generator().manual_seed(2)).cuda()
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    lo = torch.from_numpy(np.repeat(off[:-1], lens)).cuda()
    hi = torch.from_numpy(np.repeat(off[1:], lens)).cuda()
    ufo = torch.from_numpy(off).cuda()
    tiles = eng.conv_tiles(ufo, len(lens), sum((n + 127) // 128 for n in lens))
    fseg = (lo, hi, tiles, (ufo, len(lens)))
    eng.use_img_postnet = False                                             # the fused stack is the chunkable variant
    whole = eng.postnet(before, fseg, F_)
    seen = []
    chunks = planmod.output_chunks(off, 4)
    assert len(chunks) >= 3
    chunked = eng.postnet(before, fseg, F_, chunks, lambda k, out, f0, f1: seen.append((k, f0, f1)))
    eng.use_img_postnet = True
    torch.cuda.synchronize()
    assert seen == [(k, c[2], c[3]) for k, c in enumerate(chunks)]
    assert torch.equal(whole, chunked)


@pytest.mark.parametrize("kind,groups", [("S", [2, 4]), ("T", [4, 16])])
def test_decoder_group_mode_bit_identical(bf16_engines, kind, groups):
    """Group mode (several CTAs split one tile's gate columns, z exchanged through L2 with counter barriers) must
    give exactly the single-CTA-per-tile result: the arithmetic of a row does not depend on which CTA runs a chunk."""
    eng, sd, hp = bf16_engines(kind)
    xs, ds = synth.synth_batch(4, 3, fixed_len=40)               # 160 rows: 2 tiles
    pl = planmod.make_plan(xs, ds)
    d, _ = eng.upload(pl)
    hn = torch.randn(pl.n_rows, hp.eunits, generator=torch.Generator().manual_seed(7)).cuda()
    frame_off, ufo, order, totals = eng.len_reg_scan(d["dur"], d["utt_off"], pl.n_utts)
    F_ = int(pl.dur.sum())
    outs = []
    for g in [1] + groups:
        eng.force_group = g
        outs.append(eng.decoder(hn, d["dur"], frame_off, order, d["row_utt"], d["row_phone"], F_, 0.1, 0.5, 5).clone())
    eng.force_group = 0
    torch.cuda.synchronize()
    for o in outs[1:]:
        assert torch.equal(o, outs[0])


@pytest.mark.parametrize("kind,n_utts", [("S", 9), ("S", 12), ("T", 6)])
def test_decoder_pair_mode_bit_identical(bf16_engines, kind, n_utts):
    """cta_group::2 decoder (a CTA pair shares every weight stage, M = 256 MMAs) vs the single-CTA kernel:
    same arithmetic per row, so identical bits; odd and even tile counts, a partial last tile."""
    eng, sd, hp = bf16_engines(kind)
    xs, ds = synth.synth_batch(n_utts, 3, fixed_len=47)          # 423 rows -> 4 tiles, 564 -> 5 tiles, 282 -> 3 tiles
    pl = planmod.make_plan(xs, ds)
    d, _ = eng.upload(pl)
    hn = torch.randn(pl.n_rows, hp.eunits, generator=torch.Generator().manual_seed(7)).cuda()
    frame_off, ufo, order, totals = eng.len_reg_scan(d["dur"], d["utt_off"], pl.n_utts)
    F_ = int(pl.dur.sum())
    eng.force_group, eng.use_pair = 1, False
    ref = eng.decoder(hn, d["dur"], frame_off, order, d["row_utt"], d["row_phone"], F_, 0.1, 0.5, 5).clone()
    eng.use_pair = True
    got = eng.decoder(hn, d["dur"], frame_off, order, d["row_utt"], d["row_phone"], F_, 0.1, 0.5, 5).clone()
    eng.force_group, eng.use_pair = 0, None
    torch.cuda.synchronize()
    assert torch.isfinite(got).all()
    assert torch.equal(got, ref), float((got - ref).abs().max())


def test_bf16_predicted_durations_and_forced_prosody():
    """Predicted-duration path on the tensor-core engine: the pass syncs once for the frame count, output lengths
    equal the sum of the durations the kernel itself predicted; forced f0/energy are honoured."""
    from fcl_taco2_b200 import model as M
    sd = dict(weights("S", 2))
    sd["duration_predictor.linear.bias"] = torch.tensor([2.5])
    m = M.from_preset("S", seed=None, device="cpu", precision="fp16")
    m.load_state_dict(sd)
    m = m.to("cuda:0").set_prenet_dropout(rate=0.5, seed=3)
    xs, _ = synth.synth_batch(5, 44)
    res = m.inference_batch(xs, return_result=True)
    outs = res.per_utterance()
    assert all(torch.isfinite(o).all() and o.shape[1] == 80 and o.shape[0] >= len(x) for o, x in zip(outs, xs))
    assert sum(o.shape[0] for o in outs) == res.out.shape[0] == int(res.utt_frame_off[-1])
    rs = np.random.RandomState(0)
    f0 = [rs.randn(len(x)).astype(np.float32) for x in xs]
    en = [rs.randn(len(x)).astype(np.float32) for x in xs]
    ds = [np.full(len(x), 3) for x in xs]
    a = m.inference_batch(xs, durs=ds, f0s=f0, energies=en)
    b = m.inference_batch(xs, durs=ds)
    assert all(o.shape == (3 * len(x), 80) for o, x in zip(a, xs))
    assert not torch.equal(a[0], b[0])
