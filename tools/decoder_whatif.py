#!/usr/bin/env python
"""What-if timing of the cta_group::2 decoder kernels with the profiling build (lib/libfcl_taco2_prof.so, built by
`fcl_taco2_b200.build.build(prof=True)`; select it with FCL_TACO2_LIB). Each switch removes one kind of traffic / work
(results are garbage on purpose): the time that disappears tells what bounds the kernel.
usage: FCL_TACO2_LIB=fcl_taco2_b200/lib/libfcl_taco2_prof.so python tools/decoder_whatif.py [S|T] [batch]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fcl_taco2_b200 import model as M, synth, plan as planmod

kind = sys.argv[1] if len(sys.argv) > 1 else "S"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
m = M.from_preset(kind, seed=0, device="cuda:0", precision="fp16").set_prenet_dropout(rate=0.5, seed=1)
eng = m.engine()
eng.use_pair = True
xs, ds = synth.synth_batch(batch, 0)
pl = planmod.make_plan(xs, ds)
flush = torch.empty(160 << 20, dtype=torch.uint8, device="cuda:0")
NOA, NOC, NOIMG, NOPHILOX = 1 << 8, 1 << 9, 1 << 10, 1 << 11


def ring(n):
    return n << 12


def time_decoder(kern, bits, reps=6):
    eng.pair_kernel, eng.pair_inflight = kern, bits
    ts = []
    for i in range(reps + 2):
        flush.fill_(1)
        eng.stage_events = []
        eng.run(pl, 0.1, 0.5, 1)
        se, eng.stage_events = eng.stage_events, None
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(sum(a.elapsed_time(b) for n, a, b in se if n == "decoder_loop"))
    return float(np.median(ts))


SKEL, NOMATH, CNORMAL = 1 << 15, 1 << 16, 1 << 17
ALU, SLEEP = 1 << 18, 1 << 19
cases = [("baseline", 0), ("L2 prefetch of the next chunk's cell state", 1 << 21),
         ("epilogue constants from global memory (as before)", 1 << 24),
         ("roles in the last 4 warps", 1 << 20), ("skeleton epilogue", SKEL),
         ("skeleton + 800 IMAD per thread-chunk", SKEL | ALU), ("skeleton + 3200 cycles asleep per chunk", SKEL | SLEEP), ("skeleton epilogue, no A", SKEL | NOA),
         ("no cell math (TMEM ld + stores)", NOMATH), ("no cell math, no Philox", NOMATH | NOPHILOX),
         ("no cell math, no Philox, no c, no image st", NOMATH | NOPHILOX | NOC | NOIMG), ("no A copies", NOA), ("no cell-state ld/st", NOC), ("no image stores", NOIMG),
         ("no Philox", NOPHILOX), ("no A, no c, no image stores", NOA | NOC | NOIMG),
         ("all four", NOA | NOC | NOIMG | NOPHILOX)]
for kern in (sys.argv[3].split(",") if len(sys.argv) > 3 else ["v1"]):
    print(f"== {kern} ({kind} batch {batch}), decoder ms per launch (median of 6)")
    for name, bits in cases + ([("ring 5 slots", ring(5)), ("ring 3 slots", ring(3)), ("ring 4 slots", ring(4)), ("ring 3 slots, no A", ring(3) | NOA)] if kern == "v1" else []):
        print(f"   {name:32s} {time_decoder(kern, bits):.3f}")
eng.pair_kernel, eng.pair_inflight = "v1", 2
