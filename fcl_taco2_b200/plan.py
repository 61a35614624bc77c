"""Host-side batch planning: ragged packing of a list of utterances.

Pure numpy (no device work). Utterances are processed longest-first so that the
groups the BiLSTM kernel forms have similar lengths; results are returned in
the caller's order.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class BatchPlan:
    n_utts: int
    n_rows: int
    perm: np.ndarray        # (B) processing order -> caller index
    ids: np.ndarray         # (P) int64, packed in processing order
    utt_off: np.ndarray     # (B+1) int32
    row_utt: np.ndarray     # (P) int32 caller-visible utterance index (dropout key)
    row_phone: np.ndarray   # (P) int32 phoneme index within its utterance (dropout key)
    seg_lo: np.ndarray      # (P) int32 first row of the row's utterance
    seg_hi: np.ndarray      # (P) int32 one past the last row
    dur: np.ndarray | None  # (P) int32 forced durations in processing order, or None
    pitch: np.ndarray | None
    energy: np.ndarray | None


def make_plan(xs, durs=None, f0s=None, energies=None, utt_ids=None) -> BatchPlan:
    B = len(xs)
    if B == 0:
        raise ValueError("empty batch")
    lens = np.array([len(x) for x in xs], dtype=np.int64)
    if (lens <= 0).any():
        raise ValueError("zero-length utterance")
    perm = np.argsort(-lens, kind="stable")
    utt_ids = np.arange(B, dtype=np.int64) if utt_ids is None else np.asarray(utt_ids, dtype=np.int64)
    off = np.zeros(B + 1, dtype=np.int64)
    off[1:] = np.cumsum(lens[perm])
    P = int(off[-1])
    ids = np.concatenate([np.asarray(xs[i], dtype=np.int64).reshape(-1) for i in perm])
    rep = np.repeat(np.arange(B), lens[perm])
    row_phone = (np.arange(P) - off[:-1][rep]).astype(np.int32)
    row_utt = utt_ids[perm][rep].astype(np.int32)
    seg_lo = off[:-1][rep].astype(np.int32)
    seg_hi = off[1:][rep].astype(np.int32)

    def cat(vals, dtype, what):
        if vals is None:
            return None
        parts = []
        for i in perm:
            v = np.asarray(vals[i]).reshape(-1)
            if v.shape[0] != lens[i]:
                raise ValueError(f"{what}[{i}] has {v.shape[0]} entries for {lens[i]} phonemes")
            parts.append(v.astype(dtype))
        return np.concatenate(parts)

    dur = cat(durs, np.int32, "durs")
    if dur is not None and (dur < 0).any():
        raise ValueError("negative duration")
    if (f0s is None) != (energies is None):
        raise ValueError("f0 and energy must be forced together (e2e_tts_tacotron2_sa.py:649-651)")
    return BatchPlan(B, P, perm, ids, off.astype(np.int32), row_utt, row_phone, seg_lo, seg_hi, dur,
                     cat(f0s, np.float32, "f0s"), cat(energies, np.float32, "energies"))


def shard_utterances(costs, world_size: int):
    """Deal utterances to ranks so that the summed cost per rank is balanced
    (longest-processing-time greedy). -> list (per rank) of utterance indices, ascending."""
    costs = np.asarray(costs, dtype=np.int64)
    order = np.argsort(-costs, kind="stable")
    loads = np.zeros(world_size, dtype=np.int64)
    out = [[] for _ in range(world_size)]
    for i in order:
        r = int(np.argmin(loads))
        out[r].append(int(i))
        loads[r] += int(costs[i])
    return [sorted(o) for o in out]
