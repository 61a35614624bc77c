import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from fcl_taco2_b200 import hparams, pack
from fcl_taco2_b200.engine import Engine
from tests.helpers import weights
hp = hparams.preset("S"); sd = weights("S", 0)
eng = Engine(hp, pack.pack_fp32(sd, hp), "cuda:0", "fp16")
lens = [300, 5, 1, 112, 113, 64, 700]
F_ = sum(lens)
before = torch.randn(F_, 80, generator=torch.Generator().manual_seed(1)).cuda()
off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
lo = torch.from_numpy(np.repeat(off[:-1], lens)).cuda(); hi = torch.from_numpy(np.repeat(off[1:], lens)).cuda()
ufo = torch.from_numpy(off).cuda()
tiles = eng.conv_tiles(ufo, len(lens), sum((n + 127) // 128 for n in lens))
ref = eng.postnet(before, (lo, hi, tiles), F_)
fused = eng.postnet(before, (lo, hi, tiles, (ufo, len(lens))), F_)
torch.cuda.synchronize()
e = (fused - ref).abs().max(dim=1).values.cpu().numpy()
print("max", e.max(), "mean", e.mean(), "frac rows > 1e-3:", (e > 1e-3).mean())
for k in range(len(lens)):
    ee = e[off[k]:off[k+1]]
    bad = np.nonzero(ee > 1e-3)[0]
    print("utt", k, "len", lens[k], "max", ee.max(), "bad rows", bad[:20], "... n", len(bad))
# again with fresh run to check determinism
fused2 = eng.postnet(before, (lo, hi, tiles, (ufo, len(lens))), F_)
torch.cuda.synchronize()
print("fused run-to-run diff", float((fused2 - fused).abs().max()))
from oracle import restate
def ranges(idx):
    out=[]; 
    for i in idx:
        if out and i==out[-1][1]+1: out[-1][1]=i
        else: out.append([i,i])
    return out
for k in (0, 6):
    ee = e[off[k]:off[k+1]]
    print("utt", k, "bad ranges", ranges(np.nonzero(ee > 1e-4)[0].tolist()))
    o = restate.postnet(sd, before[off[k]:off[k+1]].cpu())
    ef = (fused[off[k]:off[k+1]].cpu() - o).abs().max(dim=1).values.numpy()
    er = (ref[off[k]:off[k+1]].cpu() - o).abs().max(dim=1).values.numpy()
    bad = np.nonzero(ee > 1e-3)[0]
    good = np.nonzero(ee <= 1e-4)[0]
    print("   vs oracle on bad rows: fused mean %.4f unfused mean %.4f ; on good rows: fused %.4f unfused %.4f" % (ef[bad].mean(), er[bad].mean(), ef[good].mean(), er[good].mean()))
