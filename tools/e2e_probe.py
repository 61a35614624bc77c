import sys, os, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from fcl_taco2_b200 import model as M, synth, plan as planmod
dev = torch.device("cuda", 0)
m = M.from_preset("S", seed=0, device=dev, precision="fp16").set_prenet_dropout(rate=0.5, seed=1)
xs, ds = synth.synth_batch(1024, 0)
flush = torch.empty(160 << 20, dtype=torch.uint8, device=dev)
def batches(k):
    for _ in range(k): yield {"xs": xs, "durs": ds}
def run(k, **kw):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); t0 = time.perf_counter(); e0.record()
    for outs in m.inference_stream(batches(k), before_batch=lambda: flush.fill_(1), **kw): pass
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k, (time.perf_counter() - t0) * 1e3 / k
run(6)
print("stream (D2H on):  ms/batch gpu, wall", run(20))
# host-only cost of planning
t0 = time.perf_counter()
for _ in range(20): pl = m._plan(xs, ds)
print("plan ms:", (time.perf_counter() - t0) * 1e3 / 20)
# no D2H: inference_batch loop without copies (host plans + launches only), async
eng = m.engine()
def loop_nocopy(k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(k):
        flush.fill_(1)
        pl = m._plan(xs, ds)
        res = eng.run(pl, 0.1, 0.5, 1)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
loop_nocopy(3)
print("plan+run, no D2H: ms/batch", loop_nocopy(20))
pl = m._plan(xs, ds)
def loop_noplan(k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(k):
        flush.fill_(1)
        res = eng.run(pl, 0.1, 0.5, 1)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
print("run only (upload+pass), no plan, no D2H: ms/batch", loop_noplan(20))
# host time of one enqueue (no sync)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): res = eng.run(pl, 0.1, 0.5, 1)
t1 = time.perf_counter(); torch.cuda.synchronize()
print("host enqueue ms per pass:", (t1 - t0) * 1e3 / 10)
