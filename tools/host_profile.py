"""cProfile of the host side of one pass (Engine.run): where do the ~3 ms of enqueue time go?"""
import sys, os, cProfile, pstats, io, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fcl_taco2_b200 import model as M, synth
m = M.from_preset("S", seed=0, device="cuda:0", precision="fp16").set_prenet_dropout(rate=0.5, seed=1)
xs, ds = synth.synth_batch(int(sys.argv[1]) if len(sys.argv) > 1 else 1024, 0)
pl = m._plan(xs, ds)
eng = m.engine()
for _ in range(3): eng.run(pl, 0.1, 0.5, 1)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(20): eng.run(pl, 0.1, 0.5, 1)
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
print(s.getvalue()[:6000])
