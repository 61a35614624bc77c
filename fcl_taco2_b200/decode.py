"""Batch-aware re-host of the reference's decode driver.

Mirrors `tts.decode` / `tts_distill.decode` (reference tts.py:605-686, tts_distill.py:626-718) and the flags of
`tts_decode.py:17-209` that matter at inference, without importing chainer / espnet / kaldiio:

  * model.json = [idim, odim, vars(train_args)]      (espnet get_model_conf; tts.py:341-348 writes it)
  * checkpoint: `torch.load(path)["model"]` for snapshots / amp checkpoints, else a bare state_dict
    (espnet torch_load; tts_distill.py:647-651)
  * data JSON: js["utts"][utt]["output"][0]["tokenid"] = "1 2 3 ..." (io_utils_fcl.py:142-147, preprocess.py:199-241);
    --pad-eos appends eos = shape[1] - 1 (io_utils_fcl.py:325-326)
  * output: Kaldi "ark,scp" float32 matrices (tts.py:652,674: kaldiio.WriteHelper("ark,scp:{o}.ark,{o}.scp"))
  * per-utterance "inference speed = frames / sec", mean of ratios written to <exp_name>.txt (tts.py:669-684)

New: utterances are decoded `--batch-size` at a time through `inference_stream` (the reference loops one by one);
the device->host copy and the ark write of a batch overlap the compute of the next one.

    python -m fcl_taco2_b200.decode --model exp/x/results/snapshot.ep.100 --json data/test_data.1.json \\
        --out decode/feats --ngpu 1 [--test-teacher true|false] [--batch-size 256] [--precision bf16]
"""
from __future__ import annotations

import argparse
import json
import logging
import os
import struct
import sys
import time

import numpy as np
import torch


# ----------------------------------------------------------------------------- kaldi ark/scp (binary float matrices)
class KaldiWriter:
    """Writes `<prefix>.ark` + `<prefix>.scp` like kaldiio.WriteHelper("ark,scp:...") does for float32 matrices:
    per entry  b"<key> " + b"\\0B" + b"FM " + b"\\x04" + int32 rows + b"\\x04" + int32 cols + row-major data;
    the scp line points at the byte offset of the b"\\0B" marker."""

    def __init__(self, prefix: str):
        d = os.path.dirname(prefix)
        if d:
            os.makedirs(d, exist_ok=True)
        self.ark_path = prefix + ".ark"
        self.ark = open(self.ark_path, "wb")
        self.scp = open(prefix + ".scp", "w")

    def __setitem__(self, key: str, mat: np.ndarray):
        mat = np.ascontiguousarray(mat, dtype=np.float32)
        if mat.ndim != 2:
            raise ValueError("kaldi matrices are 2-D")
        if " " in key or not key:
            raise ValueError("kaldi keys must be non-empty and contain no spaces")
        self.ark.write(key.encode() + b" ")
        off = self.ark.tell()
        self.ark.write(b"\0BFM " + b"\x04" + struct.pack("<i", mat.shape[0]) + b"\x04" + struct.pack("<i", mat.shape[1]))
        self.ark.write(mat.tobytes())
        self.scp.write(f"{key} {self.ark_path}:{off}\n")

    def close(self):
        self.ark.close()
        self.scp.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def read_kaldi_ark(path: str) -> dict:
    """Reader for the subset KaldiWriter produces (used by the tests and for spot checks)."""
    out = {}
    with open(path, "rb") as f:
        data = f.read()
    i = 0
    while i < len(data):
        j = data.index(b" ", i)
        key = data[i:j].decode()
        i = j + 1
        assert data[i:i + 5] == b"\0BFM ", "only binary float32 matrices are supported"
        i += 5
        assert data[i] == 4
        rows = struct.unpack("<i", data[i + 1:i + 5])[0]
        assert data[i + 5] == 4
        cols = struct.unpack("<i", data[i + 6:i + 10])[0]
        i += 10
        out[key] = np.frombuffer(data, dtype="<f4", count=rows * cols, offset=i).reshape(rows, cols).copy()
        i += rows * cols * 4
    return out


# ----------------------------------------------------------------------------- model.json / checkpoints / data json
def get_model_conf(model_path: str, conf_path: str | None = None):
    """-> (idim, odim, Namespace). conf defaults to `<dirname(model)>/model.json` (espnet get_model_conf)."""
    conf_path = conf_path or os.path.join(os.path.dirname(model_path), "model.json")
    with open(conf_path, "rb") as f:
        idim, odim, args = json.load(f)
    return int(idim), int(odim), argparse.Namespace(**args)


def load_weights(path: str, model) -> None:
    """Snapshot (`{'model': state_dict, ...}`: chainer torch_snapshot / amp checkpoint) or bare state_dict."""
    obj = torch.load(path, map_location="cpu")
    if isinstance(obj, dict) and "model" in obj and isinstance(obj["model"], dict):
        obj = obj["model"]
    model.load_state_dict(obj)


def read_utts(json_path: str, pad_eos: bool = False):
    """-> (utt_ids, list of int64 id arrays) in the JSON's key order (io_utils_fcl.py:142-147, :325-326)."""
    with open(json_path, "rb") as f:
        js = json.load(f)["utts"]
    ids, xs = [], []
    for utt, info in js.items():
        out0 = info["output"][0]
        x = np.fromiter(map(int, out0["tokenid"].split()), dtype=np.int64)
        if x.size == 0:
            continue                                   # the reference drops zero-length samples (io_utils_fcl.py:320-322)
        if pad_eos:
            x = np.append(x, int(out0["shape"][1]) - 1)
        ids.append(utt)
        xs.append(x)
    return ids, xs


def build_model(idim, odim, train_args, test_teacher: bool, teacher_args=None, precision="fp16"):
    from .model import Tacotron2_sa, Tacotron2_sa_student
    com = argparse.Namespace(use_fe_condition=getattr(train_args, "use_fe_condition", True),
                             append_position=getattr(train_args, "append_position", True))
    module = str(getattr(train_args, "model_module", ""))
    student = (not test_teacher) or "kd_student" in module
    if student:
        return Tacotron2_sa_student(idim, odim, train_args, com, teacher_args, precision=precision)
    return Tacotron2_sa(idim, odim, train_args, com, precision=precision)


# ----------------------------------------------------------------------------- decode
@torch.no_grad()
def decode(args, teacher_args=None):
    """Decode every utterance of `args.json` and write `<args.out>.ark/.scp`. Returns the mean of the per-utterance
    inference speeds (frames / s), the number the reference logs and writes to `<exp_name>.txt`."""
    idim, odim, train_args = get_model_conf(args.model, getattr(args, "model_conf", None))
    model = build_model(idim, odim, train_args, getattr(args, "test_teacher", True), teacher_args,
                        getattr(args, "precision", "fp16"))
    logging.info("reading model parameters from " + args.model)
    load_weights(args.model, model)
    if getattr(args, "ngpu", 1) <= 0:
        raise RuntimeError("the B200 path has no CPU fallback: use --ngpu 1")
    model = model.to(torch.device("cuda")).eval()
    utt_ids, xs = read_utts(args.json, getattr(args, "pad_eos", False))
    bs = max(1, int(getattr(args, "batch_size", 256)))
    speeds = []
    starts = list(range(0, len(xs), bs))
    batches = ({"xs": xs[b0:b0 + bs], "utt_ids": list(range(b0, min(b0 + bs, len(xs))))} for b0 in starts)
    with KaldiWriter(args.out) as writer:
        t0 = time.time()
        # batch i+1 computes while the mels of batch i travel to the host and are written to the ark
        for b0, host in zip(starts, model.inference_stream(batches)):
            dt = time.time() - t0
            t0 = time.time()
            frames = sum(h.shape[0] for h in host)
            for utt, h in zip(utt_ids[b0:b0 + bs], host):
                speeds.append(frames / dt)              # every utterance of a batch shares the batch's rate
                logging.info("inference speed = %.1f frames / sec." % speeds[-1])
                writer[utt] = h
    avg = sum(speeds) / max(len(speeds), 1)
    logging.info("average inference speed = %.1f frames / sec." % avg)
    parts = os.path.normpath(args.model).split(os.sep)
    exp_name = parts[-3] if len(parts) >= 3 else "exp"
    with open(f"{exp_name}.txt", "w") as fp:
        fp.write(str(avg))
    return avg


def _strtobool(x):
    return str(x).strip().lower() in ("y", "yes", "t", "true", "on", "1")


def get_parser():
    p = argparse.ArgumentParser(description="Synthesize mel features with the B200 FCL-taco2 path "
                                            "(flags follow the reference's tts_decode.py)")
    p.add_argument("--ngpu", default=1, type=int)
    p.add_argument("--backend", default="pytorch")
    p.add_argument("--out", type=str, required=True, help="output prefix: writes <out>.ark and <out>.scp")
    p.add_argument("--json", type=str, required=True)
    p.add_argument("--model", type=str, required=True)
    p.add_argument("--model-conf", type=str, default=None)
    p.add_argument("--pad-eos", default=False, type=_strtobool)
    p.add_argument("--test-teacher", default=True, type=_strtobool)
    p.add_argument("--teacher-conf", type=str, default=None,
                   help="teacher YAML (conf/train_pytorch_tacotron2.sa.teacher.yaml) for the student's KD tensor shapes")
    p.add_argument("--batch-size", default=256, type=int)
    p.add_argument("--precision", default="fp16", choices=["fp16", "bf16", "fp32"])
    p.add_argument("--verbose", "-V", default=0, type=int)
    # accepted for command-line compatibility; dead in the reference too (tts_decode.py:67-93, never read by inference())
    for dead, typ in (("--maxlenratio", float), ("--minlenratio", float), ("--threshold", float), ("--seed", int),
                      ("--debugmode", int), ("--use-att-constraint", _strtobool), ("--backward-window", int),
                      ("--forward-window", int), ("--save-durations", _strtobool), ("--save-focus-rates", _strtobool),
                      ("--use-amp", _strtobool)):
        p.add_argument(dead, default=None, type=typ)
    return p


def main(argv=None):
    args = get_parser().parse_args(argv)
    logging.basicConfig(level=logging.INFO if args.verbose > 0 else logging.WARN,
                        format="%(asctime)s (%(module)s:%(lineno)d) %(levelname)s: %(message)s")
    teacher_args = None
    if not args.test_teacher and args.teacher_conf:
        from .hparams import namespace_from_yaml
        teacher_args = namespace_from_yaml(args.teacher_conf)
    return decode(args, teacher_args)


if __name__ == "__main__":
    main(sys.argv[1:])
