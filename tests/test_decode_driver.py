"""Decode driver re-host: kaldi ark/scp writer, model.json / checkpoint / data-JSON handling (CPU),
and an end-to-end decode of a synthetic experiment directory on the GPU."""
import argparse
import json
import os

import numpy as np
import pytest
import torch

from fcl_taco2_b200 import decode as D, hparams, synth
from tests.helpers import weights


def _make_exp(tmp_path, kind="S", n_utts=5, snapshot=True):
    hp = hparams.preset(kind)
    exp = tmp_path / "exp" / "fcl_taco2_test" / "results"
    exp.mkdir(parents=True)
    targs = {k: v for k, v in hp.to_dict().items() if k not in ("idim", "odim")}
    targs["model_module"] = ("nets.knowledge_distillation.e2e_tts_tacotron2_sa_kd_student:Tacotron2_sa" if kind == "S"
                             else "nets.teacher_training.e2e_tts_tacotron2_sa:Tacotron2_sa")
    json.dump([hp.idim, hp.odim, targs], open(exp / "model.json", "w"))
    sd = weights(kind, 0)
    torch.save({"model": sd, "optimizer": {}} if snapshot else sd, exp / "snapshot.ep.100")
    xs, _ = synth.synth_batch(n_utts, 3)
    utts = {f"LJ{i:03d}": {"input": [], "output": [{"name": "target1", "shape": [len(x), hp.idim],
                                                     "tokenid": " ".join(map(str, x.tolist()))}], "utt2spk": "LJ"}
            for i, x in enumerate(xs)}
    json.dump({"utts": utts}, open(tmp_path / "test_data.1.json", "w"))
    return exp, xs, sd


def test_kaldi_ark_scp_round_trip(tmp_path):
    mats = {"a": np.random.rand(7, 80).astype(np.float32), "utt-2": np.random.rand(1, 80).astype(np.float32)}
    with D.KaldiWriter(str(tmp_path / "o" / "feats")) as w:
        for k, m in mats.items():
            w[k] = m
    back = D.read_kaldi_ark(str(tmp_path / "o" / "feats.ark"))
    assert list(back) == list(mats) and all(np.array_equal(back[k], mats[k]) for k in mats)
    for line in open(tmp_path / "o" / "feats.scp"):                      # scp offsets point at the \0B marker
        key, loc = line.split()
        path, off = loc.rsplit(":", 1)
        raw = open(path, "rb").read()
        assert raw[int(off):int(off) + 5] == b"\0BFM " and raw[int(off) - len(key) - 1:int(off) - 1].decode() == key
    with pytest.raises(ValueError):
        D.KaldiWriter(str(tmp_path / "x"))["bad key"] = mats["a"]


def test_model_conf_weights_and_json(tmp_path):
    exp, xs, sd = _make_exp(tmp_path)
    idim, odim, targs = D.get_model_conf(str(exp / "snapshot.ep.100"))
    assert (idim, odim, targs.dunits, targs.postnet_chans) == (76, 80, 256, 128)
    m = D.build_model(idim, odim, targs, test_teacher=False)
    D.load_weights(str(exp / "snapshot.ep.100"), m)                      # {'model': ...} snapshot, KD keys ignored
    assert torch.equal(m.state_dict()["dec.feat_out.weight"], sd["dec.feat_out.weight"])
    ids, got = D.read_utts(str(tmp_path / "test_data.1.json"))
    assert ids == [f"LJ{i:03d}" for i in range(5)] and all(np.array_equal(a, b) for a, b in zip(got, xs))
    _, padded = D.read_utts(str(tmp_path / "test_data.1.json"), pad_eos=True)
    assert all(p[-1] == 75 and len(p) == len(x) + 1 for p, x in zip(padded, xs))
    args = D.get_parser().parse_args(["--out", "o", "--json", "j", "--model", "m", "--test-teacher", "False",
                                      "--pad-eos", "False", "--maxlenratio", "10"])
    assert args.test_teacher is False and args.batch_size == 256


@pytest.mark.gpu
def test_decode_end_to_end(tmp_path, monkeypatch):
    """Random-init duration predictors emit zeros (the reference asserts on that too), so bias the duration head."""
    exp, xs, sd = _make_exp(tmp_path, snapshot=False)
    sd = dict(sd)
    sd["duration_predictor.linear.bias"] = torch.tensor([2.5])           # durations ~ exp(2.5 +- 1) - 1: all >= 1
    torch.save(sd, exp / "snapshot.ep.100")
    monkeypatch.chdir(tmp_path)
    args = D.get_parser().parse_args(["--out", str(tmp_path / "decode" / "feats"), "--json", str(tmp_path / "test_data.1.json"),
                                      "--model", str(exp / "snapshot.ep.100"), "--test-teacher", "False", "--batch-size", "2",
                                      "--precision", "fp32"])
    from fcl_taco2_b200 import model as M
    orig = M.Tacotron2_sa.__init__

    def det_init(self, *a, **k):                                         # deterministic prenet for the comparison
        orig(self, *a, **k)
        self.set_prenet_dropout(rate=0.0)
    monkeypatch.setattr(M.Tacotron2_sa, "__init__", det_init)
    avg = D.decode(args)
    assert avg > 0 and float(open(tmp_path / "fcl_taco2_test.txt").read()) == pytest.approx(avg)
    back = D.read_kaldi_ark(str(tmp_path / "decode" / "feats.ark"))
    m = M.from_preset("S", seed=None, device="cpu")
    m.load_state_dict(sd)
    m = m.to("cuda:0").set_prenet_dropout(rate=0.0)
    for i, x in enumerate(xs):
        ref = m.inference(torch.from_numpy(x), None).cpu().numpy()
        assert back[f"LJ{i:03d}"].shape == ref.shape and np.abs(back[f"LJ{i:03d}"] - ref).max() < 1e-5
