#!/usr/bin/env python
"""How long does the D2H copy of one batch's mels (185 MB, pinned) take (a) alone, (b) while the next pass computes,
and does splitting it over several streams help? (e2e of bench.py = max(pass, D2H) if the two overlap perfectly.)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fcl_taco2_b200 import model as M, synth, plan as planmod
dev = torch.device("cuda", 0)
m = M.from_preset("S", seed=0, device=dev, precision="fp16").set_prenet_dropout(rate=0.5, seed=1)
xs, ds = synth.synth_batch(1024, 0)
pl = m._plan(xs, ds)
eng = m.engine()
res = eng.run(pl, 0.1, 0.5, 1)
torch.cuda.synchronize()
out = res.out
n = out.shape[0]
host = torch.empty((n, out.shape[1]), dtype=torch.float32, pin_memory=True)
streams = [torch.cuda.Stream(dev) for _ in range(4)]

def copy(k):
    """enqueue the D2H split in k row chunks on k streams; returns (start, stop) events per chunk"""
    evs = []
    b = [n * i // k for i in range(k + 1)]
    for i in range(k):
        s = streams[i]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(s):
            e0.record(s)
            host[b[i]:b[i + 1]].copy_(out[b[i]:b[i + 1]], non_blocking=True)
            e1.record(s)
        evs.append((e0, e1))
    return evs

for k in (1, 2, 4):
    for busy in (False, True):
        ts = []
        for rep in range(5):
            torch.cuda.synchronize()
            t0 = torch.cuda.Event(enable_timing=True); t0.record()
            evs = copy(k)
            if busy:
                r2 = eng.run(pl, 0.1, 0.5, 1)
            torch.cuda.synchronize()
            ts.append(max(t0.elapsed_time(e1) for _, e1 in evs))
        gb = n * out.shape[1] * 4 / 1e9
        print(f"D2H {gb * 1e3:.0f} MB in {k} chunk(s), GPU {'running a pass' if busy else 'idle'}: {np.median(ts):.3f} ms = {gb / (np.median(ts) * 1e-3):.1f} GB/s")
# pass time with and without a concurrent copy
for busy in (False, True):
    ts = []
    for rep in range(5):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if busy:
            copy(1)
        e0.record(); r2 = eng.run(pl, 0.1, 0.5, 1); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"pass (upload + kernels) with{'' if busy else 'out'} a concurrent D2H: {np.median(ts):.3f} ms")
