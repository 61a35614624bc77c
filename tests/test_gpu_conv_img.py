"""fcl_conv_img_bf16 / fcl_pad_rows / fcl_rows_to_image (csrc/conv_img_bf16.cu) against torch on the SAME
bf16-rounded operands: isolates layout / TMA / descriptor / barrier bugs from precision. The reference ops are the
convolutions of encoder_sa.py:134-140 and variance_predictor.py:48-66,86-90 with a zero halo per utterance."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from fcl_taco2_b200 import _lib, pack
from fcl_taco2_b200._lib import dptr

pytestmark = pytest.mark.gpu
GAP = _lib.PAD_GAP


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _pad(lens):
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    B, P = len(lens), int(off[-1])
    n_tiles = (P + GAP * (B - 1) + 127) // 128
    prow_src = torch.full((n_tiles * 128,), -7, dtype=torch.int32, device="cuda")
    prow_off = torch.empty(B + 1, dtype=torch.int32, device="cuda")
    off_d = torch.from_numpy(off).cuda()
    _lib.call("fcl_pad_rows", _lib.PadRowsParams(n_utts=B, n_rows=P, gap=GAP, utt_off=dptr(off_d), n_tiles=n_tiles,
                                                 prow_src=dptr(prow_src), prow_off=dptr(prow_off)), _stream())
    return dict(off=off, P=P, B=B, n_tiles=n_tiles, rows_alloc=n_tiles * 128 + 8, prow_src=prow_src, prow_off=prow_off)


def _to_image(x, pd, gather=None, table=None):
    chans = (table if table is not None else x).shape[1]
    img = torch.full((chans // 8 * pd["rows_alloc"] * 8,), float("nan"), dtype=pack.op_dtype(), device="cuda")
    src = (table if table is not None else x).cuda().contiguous()
    g = gather.cuda() if gather is not None else None
    _lib.call("fcl_rows_to_image", _lib.RowsToImageParams(n_tiles=pd["n_tiles"], chans=chans, src=dptr(src), ld=chans,
                                                          gather=dptr(g), prow_src=dptr(pd["prow_src"]), img=dptr(img)), _stream())
    return img


def _image_rows(img, chans, pd):
    """-> (padded rows (n_tiles*128, chans) fp32 cpu, guard rows (8, chans))"""
    v = img.float().view(chans // 8, pd["rows_alloc"], 8).permute(1, 0, 2).reshape(pd["rows_alloc"], chans).cpu()
    return v[4:-4], torch.cat([v[:4], v[-4:]])


def _lens(total, rs, lo=1, hi=150):
    lens = []
    while sum(lens) < total:
        lens.append(int(min(rs.randint(lo, hi), total - sum(lens))))
    return lens


def _ref_conv(a, w, bias, lens, taps, relu=True):
    ab, wb = a.to(pack.op_dtype()).float(), w.to(pack.op_dtype()).float()
    outs, o = [], 0
    for n in lens:
        y = F.conv1d(ab[o:o + n].t().unsqueeze(0), wb.permute(2, 1, 0).contiguous(), bias, 1, taps // 2)[0].t()
        outs.append(y)
        o += n
    y = torch.cat(outs)
    return torch.relu(y) if relu else y


def test_pad_rows_and_rows_to_image():
    rs = np.random.RandomState(0)
    lens = [1, 5, 130, 1, 64, 127, 3] + _lens(900, rs)
    pd = _pad(lens)
    src = pd["prow_src"].cpu().numpy()
    off = pd["off"]
    want = np.full(pd["n_tiles"] * 128, -1, dtype=np.int64)
    for u, n in enumerate(lens):
        want[off[u] + GAP * u: off[u] + GAP * u + n] = np.arange(off[u], off[u] + n)
    assert np.array_equal(src, want)
    assert np.array_equal(pd["prow_off"].cpu().numpy(), off + GAP * np.arange(len(lens) + 1))
    x = torch.randn(pd["P"], 64, generator=torch.Generator().manual_seed(0))
    rows, guard = _image_rows(_to_image(x, pd), 64, pd)
    assert (guard == 0).all()
    valid = torch.from_numpy(want >= 0)
    assert (rows[~valid] == 0).all()
    assert torch.equal(rows[valid], x.to(pack.op_dtype()).float())
    # embedding gather
    table = torch.randn(76, 64, generator=torch.Generator().manual_seed(1))
    ids = torch.randint(0, 76, (pd["P"],), generator=torch.Generator().manual_seed(2))
    rows, _ = _image_rows(_to_image(None, pd, gather=ids, table=table), 64, pd)
    assert torch.equal(rows[valid], table[ids].to(pack.op_dtype()).float())


@pytest.mark.parametrize("total,cin,cout,taps", [
    (700, 256, 256, 5),          # S encoder conv; 6 tiles
    (90, 256, 256, 5),           # one tile (the peer CTA of the pair gets a tile past the end)
    (129 * 128, 256, 256, 5),    # ~131 tiles -> 66 super-tiles on 74 pairs
    (400 * 128, 256, 256, 5),    # ~406 tiles: several super-tiles per pair, TMEM double buffering across tiles
    (1500, 512, 512, 5),         # T encoder conv: two N blocks re-read the resident window (8 A stages)
    (1000, 384, 128, 3),         # nb = 128
    (1000, 64, 192, 1),          # one K stage, nb = 192, plain linear
    (600 * 128, 128, 128, 5),    # postnet-sized layer: RESIDENT weights (10 stages), ~4 super-tiles per pair, 6 window stages
    (1000, 128, 128, 5),         # resident weights, fewer tiles than CTA pairs
    (90, 128, 128, 3),           # resident weights, one tile (the peer's tile is past the end)
    (3000, 64, 128, 5),          # resident weights, one K chunk per tile (3 window stages)
])
def test_conv_img_image_epilogue(total, cin, cout, taps):
    rs = np.random.RandomState(total + cin)
    g = torch.Generator().manual_seed(total + cout)
    lens = _lens(total, rs)
    pd = _pad(lens)
    a = torch.randn(pd["P"], cin, generator=g)
    w = torch.randn(taps, cin, cout, generator=g) / np.sqrt(cin * taps)
    bias = torch.randn(cout, generator=g) * 0.1
    wp, nb = pack.pack_conv_pair(w)
    wp, bias_d = wp.cuda(), bias.cuda()
    img = _to_image(a, pd)
    out = torch.full((cout // 8 * pd["rows_alloc"] * 8,), float("nan"), dtype=pack.op_dtype(), device="cuda")
    _lib.call("fcl_conv_img_bf16", _lib.ConvImgParams(n_tiles=pd["n_tiles"], cin=cin, cout=cout, taps=taps, nb=nb,
                                                      act=_lib.ACT_RELU, epi=_lib.EPI_IMAGE, in_img=dptr(img), w_packed=dptr(wp),
                                                      bias=dptr(bias_d), prow_src=dptr(pd["prow_src"]), out_img=dptr(out)), _stream())
    torch.cuda.synchronize()
    rows, guard = _image_rows(out, cout, pd)
    assert torch.isfinite(rows).all() and (guard == 0).all()
    valid = (pd["prow_src"].cpu() >= 0)
    assert (rows[~valid] == 0).all(), "gap rows of the output image must be zero (they are the next layer's halo)"
    ref = _ref_conv(a, w, bias, lens, taps)
    err = float((rows[valid] - ref).abs().max())
    assert err < 1e-2 * max(1.0, float(ref.abs().max())), err        # output rounded to bf16: 2^-9 relative


@pytest.mark.parametrize("total,cin,C,head", [(900, 256, 384, False), (900, 384, 384, True), (300 * 128, 256, 384, True),
                                              (700, 512, 384, False)])
def test_conv_img_layernorm_epilogues(total, cin, C, head):
    """Conv k3 + bias -> ReLU -> LayerNorm(C, eps 1e-12) -> image, or -> Linear(C, 1) (+ duration rounding)."""
    rs = np.random.RandomState(total + cin + head)
    g = torch.Generator().manual_seed(cin + C + head)
    lens = _lens(total, rs)
    pd = _pad(lens)
    a = torch.randn(pd["P"], cin, generator=g)
    w = torch.randn(3, cin, C, generator=g) / np.sqrt(cin * 3)
    bias, gamma, beta = torch.randn(C, generator=g) * 0.1, torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.1
    hw, hb = torch.randn(C, generator=g) / np.sqrt(C), 1.3
    wp, nb = pack.pack_conv_pair(w)
    dev = [t.cuda() for t in (wp, bias, gamma, beta, hw)]
    img = _to_image(a, pd)
    out = torch.full((C // 8 * pd["rows_alloc"] * 8,), float("nan"), dtype=pack.op_dtype(), device="cuda")
    head_out = torch.full((pd["P"],), float("nan"), device="cuda")
    dur_out = torch.full((pd["P"],), -1, dtype=torch.int32, device="cuda")
    _lib.call("fcl_conv_img_bf16", _lib.ConvImgParams(
        n_tiles=pd["n_tiles"], cin=cin, cout=C, taps=3, nb=nb, act=_lib.ACT_RELU,
        epi=_lib.EPI_LN_HEAD if head else _lib.EPI_LN_IMAGE, in_img=dptr(img), w_packed=dptr(dev[0]), bias=dptr(dev[1]),
        prow_src=dptr(pd["prow_src"]), out_img=dptr(out), gamma=dptr(dev[2]), beta=dptr(dev[3]), head_w=dptr(dev[4]) if head else None,
        head_b=hb, head_out=dptr(head_out) if head else None, dur_out=dptr(dur_out) if head else None), _stream())
    torch.cuda.synchronize()
    y = F.layer_norm(_ref_conv(a, w, bias, lens, 3), (C,), gamma, beta, 1e-12)
    if head:
        ref = y @ hw + hb
        got = head_out.cpu()
        assert torch.isfinite(got).all()
        assert float((got - ref).abs().max()) < 5e-3 * max(1.0, float(ref.abs().max()))
        want_d = torch.clamp(torch.round(got.exp() - 1.0), 0, 1023).to(torch.int32)        # rounding of the kernel's own head
        assert torch.equal(dur_out.cpu(), want_d)
    else:
        rows, guard = _image_rows(out, C, pd)
        valid = (pd["prow_src"].cpu() >= 0)
        assert (guard == 0).all() and (rows[~valid] == 0).all()
        assert float((rows[valid] - y).abs().max()) < 1e-2 * max(1.0, float(y.abs().max()))


@pytest.mark.parametrize("half", [False, True])
@pytest.mark.parametrize("total,cin,cout", [(900, 256, 1024), (200 * 128, 256, 1024), (600, 512, 2048)])
def test_conv_img_blocked_epilogues(total, cin, cout, half):
    """The BiLSTM input projection: linear + bias -> fp32, column-blocked [cout/16][n_tiles*128][16] by padded row."""
    rs = np.random.RandomState(total + cout)
    g = torch.Generator().manual_seed(total)
    lens = _lens(total, rs)
    pd = _pad(lens)
    a = torch.randn(pd["P"], cin, generator=g)
    w = torch.randn(1, cin, cout, generator=g) / np.sqrt(cin)
    bias = torch.randn(cout, generator=g) * 0.1
    wp, nb = pack.pack_conv_pair(w)
    wp, bias_d = wp.cuda(), bias.cuda()
    img = _to_image(a, pd)
    R = pd["n_tiles"] * 128
    out = torch.full((cout // 16 * R * 16,), float("nan"), device="cuda", dtype=torch.float16 if half else torch.float32)
    _lib.call("fcl_conv_img_bf16", _lib.ConvImgParams(n_tiles=pd["n_tiles"], cin=cin, cout=cout, taps=1, nb=nb, act=_lib.ACT_NONE,
                                                      epi=_lib.EPI_BLOCKED_F16 if half else _lib.EPI_BLOCKED_F32, in_img=dptr(img), w_packed=dptr(wp), bias=dptr(bias_d),
                                                      prow_src=dptr(pd["prow_src"]), out_blk=dptr(out)), _stream())
    torch.cuda.synchronize()
    rows = out.float().view(cout // 16, R, 16).permute(1, 0, 2).reshape(R, cout).cpu()
    valid = (pd["prow_src"].cpu() >= 0)
    ref = a.to(pack.op_dtype()).float() @ w[0].to(pack.op_dtype()).float() + bias
    got = rows[valid]
    assert torch.isfinite(got).all()
    assert float((got - ref).abs().max()) < (3e-3 if half else 2e-3) * max(1.0, float(ref.abs().max()))
