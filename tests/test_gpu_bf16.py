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
    ab = a.to(torch.bfloat16).float()
    wb = w.to(torch.bfloat16).float()
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
    p = _lib.ConvGemmBf16Params(rows=rows, cin=cin, cout=cout, taps=taps, a=dptr(a_dev), lda=cin, gather=dptr(gather),
                                seg_lo=dptr(lo) if taps > 1 else None, seg_hi=dptr(hi) if taps > 1 else None,
                                w_packed=dptr(wp), ntile=ntile, kstage=kstage, bias=dptr(bias_d), residual=dptr(res_d),
                                ldr=cout, out=dptr(out), ldo=cout, act=act)
    _lib.call("fcl_conv_gemm_bf16", p, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ref = _ref_conv(a_ref, w, bias, seg_lens if taps > 1 else [rows], taps, act, res)
    got = out.cpu()
    assert torch.isfinite(got).all()
    err = float((got - ref).abs().max())
    assert err < 2e-3 * max(1.0, float(ref.abs().max())), err
