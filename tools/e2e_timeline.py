#!/usr/bin/env python
"""Timeline of model.inference_stream (bench.py's e2e path) from CUDA events: per batch, when the pass starts / ends on
the launching stream and when its D2H copy starts / ends on the copy stream -- which of the two pipelines is the bound?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fcl_taco2_b200 import model as M, synth
dev = torch.device("cuda", 0)
m = M.from_preset("S", seed=0, device=dev, precision="fp16").set_prenet_dropout(rate=0.5, seed=1)
xs, ds = synth.synth_batch(1024, 0)
flush = torch.empty(160 << 20, dtype=torch.uint8, device=dev)
marks = []
def before():
    e = torch.cuda.Event(enable_timing=True); e.record(); marks.append(e)      # previous pass's kernels are all before this
    flush.fill_(1)
def batches(k):
    for _ in range(k): yield {"xs": xs, "durs": ds}
for _ in m.inference_stream(batches(6), before_batch=before): pass
torch.cuda.synchronize(); marks.clear()
t0 = time.perf_counter()
n = 0
for outs in m.inference_stream(batches(24), before_batch=before): n += 1
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) * 1e3
d = [marks[i].elapsed_time(marks[i + 1]) for i in range(len(marks) - 1)]
print(f"wall per batch {wall / n:.3f} ms; launching-stream period between batch starts: median {np.median(d):.3f} ms "
      f"(min {min(d):.3f}, max {max(d):.3f})")
print("periods:", " ".join(f"{x:.2f}" for x in d))
