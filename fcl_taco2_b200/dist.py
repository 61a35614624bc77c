"""Multi-GPU: utterance sharding (no collective on the hot path) + one final gather of mels.

One process per GPU (torchrun). Every rank holds a full weight copy and runs the whole
pass on its own utterances; the only communication is the gather of the ragged mel
buffers to rank 0 (ncclSend/ncclRecv pairs through torch.distributed P2P ops) preceded
by a tiny all_gather of frame counts. Works with the gloo backend for CPU tests.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .plan import shard_utterances


def my_shard(costs, rank: int | None = None, world_size: int | None = None):
    """Indices of the utterances this rank processes (balanced by `costs`, e.g. phoneme or frame counts)."""
    rank = dist.get_rank() if rank is None else rank
    world_size = dist.get_world_size() if world_size is None else world_size
    return shard_utterances(costs, world_size)[rank]


def exchange_counts(n_frames: int, device, group=None):
    """Frame count of every rank (world_size * 8 bytes). With forced durations the count is known on the host
    before any kernel runs, so this tiny collective can be issued at the START of a pass."""
    ws = dist.get_world_size(group)
    cnt = torch.tensor([n_frames], dtype=torch.int64, device=device)
    cnts = [torch.empty_like(cnt) for _ in range(ws)]
    dist.all_gather(cnts, cnt, group=group)
    return [int(c.item()) for c in cnts]


def gather_mels(out: torch.Tensor, dst: int = 0, group=None, counts=None):
    """out (F_rank, odim) on every rank -> on `dst`: (list of per-rank tensors); elsewhere None.
    Message sizes are data dependent: `counts` (from exchange_counts) avoids a host sync at the end of the pass."""
    ws, rank = dist.get_world_size(group), dist.get_rank(group)
    odim = out.shape[1]
    if counts is None:
        counts = exchange_counts(out.shape[0], out.device, group)
    if ws == 1:
        return [out]
    if rank == dst:
        bufs = [out if r == dst else torch.empty((counts[r], odim), dtype=out.dtype, device=out.device)
                for r in range(ws)]
        ops = [dist.P2POp(dist.irecv, bufs[r], r, group) for r in range(ws) if r != dst and counts[r] > 0]
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return bufs
    if out.shape[0] > 0:
        for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, out.contiguous(), dst, group)]):
            req.wait()
    return None


def scatter_results(shards, per_rank_outputs, n_total):
    """Rank-0 helper: put per-rank per-utterance outputs back into the caller's global order."""
    res = [None] * n_total
    for idxs, outs in zip(shards, per_rank_outputs):
        for i, o in zip(idxs, outs):
            res[i] = o
    return res
