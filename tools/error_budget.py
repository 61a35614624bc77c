import sys; sys.path.insert(0,'/root/repo')
import torch, numpy as np
from fcl_taco2_b200 import hparams, pack, plan as planmod, synth
from fcl_taco2_b200.engine import Engine
from tests.helpers import load_golden, weights, err
g = load_golden("S_n500_stress")
hp = hparams.preset("S"); sd = weights("S", g["weight_seed"]); pk = pack.pack_fp32(sd, hp)
front = ["enc_conv0","enc_conv1","enc_conv2","blstm_wih"]
pred = ["dur_conv0","dur_conv1","pitch_conv0","pitch_conv1","energy_conv0","energy_conv1"]
post = ["post_conv%d"%i for i in range(5)]
cfgs = {"fp32": ("fp32", None, False), "dec only": ("fp16", [], True), "enc only": ("fp16", front, False),
        "pred only": ("fp16", pred, False), "post only": ("fp16", post, False), "all": ("fp16", None, True),
        "all but pred": ("fp16", front+post, True)}
pl = planmod.make_plan([g["x"]], [g["dur"]])
for name,(prec,keys,dec) in cfgs.items():
    eng = Engine(hp, pk, "cuda:0", prec, bf16_gemms=keys, bf16_decoder=dec) if prec=="fp16" else Engine(hp, pk, "cuda:0")
    res = eng.run(pl, 0.1, 0.0, 0)
    print(f"{name:14s} max-abs %.3e mean-L1 %.3e" % err(res.out.cpu(), g["out"]))
