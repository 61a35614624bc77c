#!/usr/bin/env python
"""In-kernel timeline of fcl_conv_img_bf16 (CTA 0): where a tile's time goes (MMA issue vs epilogue passes).
usage: python tools/conv_img_trace.py [ln_image|ln_head|image|blocked] [tiles]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fcl_taco2_b200 import _lib, pack
from fcl_taco2_b200._lib import dptr

kind = sys.argv[1] if len(sys.argv) > 1 else "ln_image"
tiles = int(sys.argv[2]) if len(sys.argv) > 2 else 657
cfg = {"ln_image": (256, 384, 3, _lib.EPI_LN_IMAGE, _lib.ACT_RELU), "ln_head": (384, 384, 3, _lib.EPI_LN_HEAD, _lib.ACT_RELU),
       "image": (256, 256, 5, _lib.EPI_IMAGE, _lib.ACT_RELU), "blocked": (256, 1024, 1, _lib.EPI_BLOCKED_F16, _lib.ACT_NONE)}[kind]
cin, cout, taps, epi, act = cfg
if len(sys.argv) > 5:
    cin, cout, taps = int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
force_nb = int(sys.argv[6]) if len(sys.argv) > 6 else None
g = torch.Generator().manual_seed(0)
w = torch.randn(taps, cin, cout, generator=g) / np.sqrt(cin * taps)
wp, nb = pack.pack_conv_pair(w, force_nb)
wp = wp.cuda()
rows_alloc = tiles * 128 + 8
img = (torch.randn(cin // 8 * rows_alloc * 8, generator=g) * 0.5).to(pack.op_dtype()).cuda()
prow_src = torch.arange(tiles * 128, dtype=torch.int32).cuda()
out_img = torch.empty(cout // 8 * rows_alloc * 8, dtype=pack.op_dtype(), device="cuda")
out_blk = torch.empty(cout // 16 * tiles * 128 * 16, dtype=torch.float16, device="cuda")
vec = lambda: torch.randn(cout, generator=g).cuda()
bias, gamma, beta, hw = vec(), vec(), vec(), vec()
head = torch.empty(tiles * 128, device="cuda")
trace = torch.zeros(2 + 2 * (4096 + 512), dtype=torch.int64, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
for it in range(3):
    trace.zero_()
    p = _lib.ConvImgParams(n_tiles=tiles, cin=cin, cout=cout, taps=taps, nb=nb, act=act, epi=epi, in_img=dptr(img), w_packed=dptr(wp),
                           bias=dptr(bias), prow_src=dptr(prow_src), out_img=dptr(out_img), out_blk=dptr(out_blk), gamma=dptr(gamma),
                           beta=dptr(beta), head_w=dptr(hw), head_b=0.1, head_out=dptr(head), trace=dptr(trace) if it == 2 else None,
                           trace_cap=4096)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); _lib.call("fcl_conv_img_bf16", p, stream); e1.record()
    torch.cuda.synchronize()
pp = _lib.ConvImgParams(n_tiles=tiles, cin=cin, cout=cout, taps=taps, nb=nb, act=act, epi=epi, in_img=dptr(img), w_packed=dptr(wp),
                        bias=dptr(bias), prow_src=dptr(prow_src), out_img=dptr(out_img), out_blk=dptr(out_blk), gamma=dptr(gamma),
                        beta=dptr(beta), head_w=dptr(hw), head_b=0.1, head_out=dptr(head), trace=None, trace_cap=0)
f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
f0.record()
for _ in range(20):
    _lib.call("fcl_conv_img_bf16", pp, stream)
f1.record()
torch.cuda.synchronize()
print(f"untraced, 20 launches back to back: {f0.elapsed_time(f1) * 1e3 / 20:.1f} us per launch")
print(f"{kind}: {tiles} tiles, cin {cin} cout {cout} taps {taps} nb {nb}: {e0.elapsed_time(e1) * 1e3:.1f} us (last launch, traced)")
t = trace.cpu().numpy()
n = int(min(t[0], 4096))
rec = sorted((int(t[3 + 2 * i]), int(t[2 + 2 * i])) for i in range(n))
t0 = rec[0][0]
try:
    for clk, ev in rec[:60]:
        print(f"{clk - t0:9d}  {ev}")
    print("   ...")
    for clk, ev in rec[-12:]:
        print(f"{clk - t0:9d}  {ev}")
    life = t[2 + 2 * 4096: 2 + 2 * (4096 + 148)].reshape(-1, 2)
    life = life[life[:, 0] > 0]
    g0 = life[:, 0].min()
    print(f"CTA life spans (globaltimer, us): first entry 0, last entry {(life[:, 0].max() - g0) / 1e3:.1f}, first exit {(life[:, 1].min() - g0) / 1e3:.1f}, "
          f"last exit {(life[:, 1].max() - g0) / 1e3:.1f}; median life {np.median(life[:, 1] - life[:, 0]) / 1e3:.1f}")
    lives = (life[:, 1] - life[:, 0]) / 1e3
    order = np.argsort(lives)
    print("slowest CTAs (blockIdx: us): " + ", ".join(f"{i}: {lives[i]:.1f}" for i in order[-16:]))
    print("life histogram (us): " + str(np.histogram(lives, bins=8)))
    issued = [c for c, e in rec if e == 200]
    if len(issued) > 4:
        d = np.diff(issued)
        print(f"records {n}; tiles of CTA 0: {len(issued)}; cycles between 'all MMAs of a tile issued': median {np.median(d):.0f}, mean {np.mean(d):.0f}, max {d.max()}")
except BrokenPipeError:
    pass
