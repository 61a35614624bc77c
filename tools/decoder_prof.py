#!/usr/bin/env python
"""Counter-mode profile of the cta_group::2 decoder kernels (CTA 0): where do the cycles of a tile-step go?
usage: python tools/decoder_prof.py [S|T] [batch] [kernels: v1,v2]
The kernels accumulate %clock deltas of the single-thread roles and one epilogue thread in shared memory and dump 48
counters at the end (layout: csrc/decoder_bf16_pair_v1.cu, DB_PROF)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fcl_taco2_b200 import model as M, synth, plan as planmod

kind = sys.argv[1] if len(sys.argv) > 1 else "S"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
kernels = (sys.argv[3] if len(sys.argv) > 3 else "v1").split(",")
m = M.from_preset(kind, seed=0, device="cuda:0", precision="fp16").set_prenet_dropout(rate=0.5, seed=1)
eng = m.engine()
eng.use_pair = True
xs, ds = synth.synth_batch(batch, 0)
pl = planmod.make_plan(xs, ds)
PH = ["P1", "L0", "L1", "FP"]
for kern in kernels:
    eng.pair_kernel = kern
    eng.dec_trace, eng.dec_prof = None, False
    for _ in range(2):
        eng.run(pl, 0.1, 0.5, 1)
    eng.dec_trace = torch.zeros(64, dtype=torch.int64, device="cuda:0")
    eng.dec_prof = True
    eng.run(pl, 0.1, 0.5, 1)
    torch.cuda.synchronize()
    c = eng.dec_trace.cpu().numpy().astype(np.float64)
    steps = max(c[32], 1)
    print(f"== {kern}: CTA 0 walked {int(steps)} tile-steps, {c[33] / steps:,.0f} cycles per tile-step (issuer total {c[33]:,.0f})")
    print("   MMA issuer, cycles per tile-step:   wait acc | wait 1st item | wait items | issue | items")
    for ph in range(4):
        v = c[4 * ph:4 * ph + 4] / steps
        print(f"     {PH[ph]}: {v[0]:9,.0f} {v[1]:9,.0f} {v[2]:9,.0f} {v[3]:9,.0f}   {c[34 + ph] / steps:6.1f}")
    tot = c[:16].reshape(4, 4).sum(0) / steps
    print(f"     sum {tot[0]:9,.0f} {tot[1]:9,.0f} {tot[2]:9,.0f} {tot[3]:9,.0f}")
    print("   epilogue thread, cycles per tile-step: wait acc | body")
    for ph in range(4):
        print(f"     {PH[ph]}: {c[16 + 2 * ph] / steps:9,.0f} {c[17 + 2 * ph] / steps:9,.0f}")
    print(f"     FP chunk 1 (composed prenet.0): {c[44] / steps:9,.0f} {c[45] / steps:9,.0f}")
    print("   producer, cycles per tile-step: wait a_ready | wait free slot")
    for ph in range(4):
        print(f"     {PH[ph]}: {c[24 + ph] / steps:9,.0f} {c[28 + ph] / steps:9,.0f}")
    print(f"   residue: epilogue {c[46] / steps:,.0f}, producer {c[47] / steps:,.0f}")
eng.dec_trace, eng.dec_prof = None, False
