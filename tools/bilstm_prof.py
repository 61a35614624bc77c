#!/usr/bin/env python
"""Counter-mode profile of fcl_bilstm_bf16 (profiling build, FCL_TACO2_LIB=.../libfcl_taco2_prof.so): where does a
step of the recurrence go? CTA (0, 0) = the longest tile, forward direction."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fcl_taco2_b200 import _lib, hparams, pack, synth
from fcl_taco2_b200._lib import dptr
from tests.helpers import weights
hp = hparams.preset("S")
packed = pack.pack_fp32(weights("S", 0), hp)
hd, E = hp.eunits // 2, hp.eunits
xs, _ = synth.synth_batch(1024, 0)
lens = np.sort(np.array([len(x) for x in xs]))[::-1].copy()
off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
P, n = int(off[-1]), len(lens)
gx = (torch.randn(P, hp.econv_chans) @ packed["blstm_wih"][0] + packed["blstm_b"]).to(pack.op_dtype()).cuda()
whh = pack.pack_bilstm_whh_bf16(packed).cuda()
c_ws = torch.zeros(((n + 31) // 32) * 2 * hd * 128, dtype=torch.float32, device="cuda")
out = torch.empty(P, E, device="cuda")
off_d = torch.from_numpy(off).cuda()
p = _lib.BiLstmBf16Params(n_utts=n, hidden=hd, tile_utts=32, utt_off=dptr(off_d), gx=dptr(gx), whh_packed=dptr(whh),
                          c_ws=dptr(c_ws), out=dptr(out))
st = torch.cuda.current_stream().cuda_stream
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    e0.record(); _lib.call("fcl_bilstm_bf16", p, st); e1.record()
torch.cuda.synchronize()
c = c_ws[:8].view(torch.int32).cpu().numpy().astype(np.float64)
steps = max(c[0], 1)
print(f"bilstm S batch 1024: {e0.elapsed_time(e1) * 1e3:.1f} us; tile 0: {int(steps)} steps")
print(f"  issuer per step: waiting for h_ready {c[1] / steps:.0f}, issuing {c[2] / steps:.0f} cycles")
for h in (0, 1):
    print(f"  epilogue half {h} per step: waiting for its accumulator {c[3 + 2 * h] / steps:.0f}, body (to h_ready arrive) {c[4 + 2 * h] / steps:.0f}")
