"""cProfile of the host side of model.inference_stream (the e2e path of bench.py): is the host the longer pole?"""
import sys, os, cProfile, pstats, io, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fcl_taco2_b200 import model as M, synth
m = M.from_preset("S", seed=0, device="cuda:0", precision="fp16").set_prenet_dropout(rate=0.5, seed=1)
xs, ds = synth.synth_batch(1024, 0)
flush = torch.empty(160 << 20, dtype=torch.uint8, device="cuda:0")
def batches(k):
    for _ in range(k): yield {"xs": xs, "durs": ds}
for _ in m.inference_stream(batches(6), before_batch=lambda: flush.fill_(1)): pass
torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
for _ in m.inference_stream(batches(30), before_batch=lambda: flush.fill_(1)): pass
pr.disable()
torch.cuda.synchronize()
print("wall ms per batch:", (time.perf_counter() - t0) * 1e3 / 30)
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22)
print(s.getvalue()[:5000])
