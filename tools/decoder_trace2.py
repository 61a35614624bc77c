#!/usr/bin/env python
"""Timeline of the cta_group::2 decoder's CTA 0 with 1 or 2 super-tiles in flight.
usage: python tools/decoder_trace2.py [S|T] [batch] [inflight] [lines]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fcl_taco2_b200 import model as M, synth, plan as planmod

kind = sys.argv[1] if len(sys.argv) > 1 else "S"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
infl = int(sys.argv[3]) if len(sys.argv) > 3 else 2
NLINES = int(sys.argv[4]) if len(sys.argv) > 4 else 120
m = M.from_preset(kind, seed=0, device="cuda:0", precision="fp16").set_prenet_dropout(rate=0.5, seed=1)
eng = m.engine()
eng.use_pair, eng.pair_inflight = True, infl
xs, ds = synth.synth_batch(batch, 0)
pl = planmod.make_plan(xs, ds)
for _ in range(2):
    eng.run(pl, 0.1, 0.5, 1)
eng.dec_trace = torch.zeros(2 + 2 * 40000, dtype=torch.int64, device="cuda:0")
eng.run(pl, 0.1, 0.5, 1)
torch.cuda.synchronize()
t = eng.dec_trace.cpu().numpy()
n = int(t[0]); rec = t[2:2 + 2 * min(n, 40000)].reshape(-1, 2)
rec = rec[np.argsort(rec[:, 1], kind="stable")]
t0 = rec[0, 1]
print("records", n, "inflight", infl)
names = {1: "mma:acc_free", 3: "mma:issued", 4: "epi:acc_ready", 5: "epi:done", 6: "prod:item"}
skip = 400                                                # past the start-up
for ev, clk in rec[skip:skip + NLINES]:
    slot, e = ev // 1000, ev % 1000
    k, rest = e // 100, e % 100
    print(f"{clk - t0:9d}  s{slot} {names.get(k, k):14s} phase {rest // 10 if k != 6 else rest} chunk {rest % 10 if k != 6 else '-'}")
# aggregates per (phase, chunk), both slots together
def span(a, b):
    out, ia = {}, {}
    for e, c in rec:
        slot, x = e // 1000, e % 1000
        if x // 100 == a: ia[(slot, x % 100)] = c
        elif x // 100 == b and (slot, x % 100) in ia:
            out.setdefault(x % 100, []).append(c - ia.pop((slot, x % 100)))
    return {k: round(float(np.mean(v))) for k, v in sorted(out.items())}
print("mma: acc free -> chunk issued:", span(1, 3))
print("epilogue: acc ready -> done  :", span(4, 5))
print("issued -> epilogue woke      :", span(3, 4))
e = rec[:, 0] % 1000
clk = rec[:, 1]
p1 = clk[(e == 400)]
print("mean cycles between P1 epilogues (= per tile-step when inflight 1, per half round when 2):", float(np.mean(np.diff(p1))) if len(p1) > 2 else None)
