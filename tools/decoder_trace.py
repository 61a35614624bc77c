#!/usr/bin/env python
"""Timeline of the tensor-core decoder's CTA 0 (debug aid): where do the cycles of a step go?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fcl_taco2_b200 import model as M, synth, plan as planmod

kind = sys.argv[1] if len(sys.argv) > 1 else "S"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
m = M.from_preset(kind, seed=0, device="cuda:0", precision="fp16").set_prenet_dropout(rate=0.5, seed=1)
eng = m.engine()
eng.use_pair = len(sys.argv) > 3 and sys.argv[3] == "pair"
NLINES = int(sys.argv[4]) if len(sys.argv) > 4 else 140
xs, ds = synth.synth_batch(batch, 0)
pl = planmod.make_plan(xs, ds)
for _ in range(2):
    eng.run(pl, 0.1, 0.5, 1)
eng.dec_trace = torch.zeros(2 + 2 * 20000, dtype=torch.int64, device="cuda:0")
eng.run(pl, 0.1, 0.5, 1)
torch.cuda.synchronize()
t = eng.dec_trace.cpu().numpy()
n = int(t[0]); rec = t[2:2 + 2 * min(n, 20000)].reshape(-1, 2)
rec = rec[np.argsort(rec[:, 1], kind="stable")]
t0 = rec[0, 1]
print("records", n)
# first ~2 steps as a timeline
names = {7: "epi:detail", 1: "mma:acc_free", 2: "mma:stage0_landed", 3: "mma:issued", 4: "epi:acc_ready", 5: "epi:done", 6: "prod:operand_ready"}
for ev, clk in rec[:NLINES]:
    kind_, rest = ev // 100, ev % 100
    print(f"{clk - t0:9d}  {names[kind_]:20s} phase {rest // 10 if kind_ != 6 else rest} chunk {rest % 10 if kind_ != 6 else '-'}")
# aggregate: per (phase, chunk): issue time (300-200), wait for stage0 (200-100), epilogue (500-400), and step length
ev = rec[:, 0]; clk = rec[:, 1]
steps = clk[ev == 600]
if len(steps) > 2:
    print("mean cycles per step:", float(np.mean(np.diff(steps))), "steps traced:", len(steps))
def pair(a, b):
    out = {}
    ia = {}
    for e, c in rec:
        if e // 100 == a: ia[e % 100] = c
        elif e // 100 == b and (e % 100) in ia:
            out.setdefault(e % 100, []).append(c - ia.pop(e % 100))
    return {k: float(np.mean(v)) for k, v in sorted(out.items())}
print("mma issue span per chunk (first stage landed -> issued):", pair(2, 3))
print("mma wait for first stage (acc free -> stage landed):", pair(1, 2))
print("epilogue per chunk (acc ready -> done):", pair(4, 5))
print("commit->epilogue wake (issued -> acc ready):", pair(3, 4))
