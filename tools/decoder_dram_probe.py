#!/usr/bin/env python
"""Does the decoder's DRAM traffic come from L2 capacity? Run the pair decoder with fewer CTA pairs (smaller scratch
working set: 640 KB per CTA) under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum`.
usage: ncu ... python tools/decoder_dram_probe.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fcl_taco2_b200 import model as M, synth, plan as planmod
m = M.from_preset("S", seed=0, device="cuda:0", precision="fp16").set_prenet_dropout(rate=0.5, seed=1)
eng = m.engine()
xs, ds = synth.synth_batch(1024, 0)
pl = planmod.make_plan(xs, ds)
for mp in [int(a) if a != "all" else None for a in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["all", "37", "18", "9"])]:
    eng.max_pairs = mp
    for _ in range(2):
        eng.run(pl, 0.1, 0.5, 1)
    torch.cuda.synchronize()
eng.max_pairs = None
