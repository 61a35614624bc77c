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


def gather_mels(out: torch.Tensor, dst: int = 0, group=None):
    """out (F_rank, odim) on every rank -> on `dst`: (list of per-rank tensors); elsewhere None.
    Message sizes are data dependent, so counts are exchanged first (world_size * 8 bytes)."""
    ws, rank = dist.get_world_size(group), dist.get_rank(group)
    odim = out.shape[1]
    cnt = torch.tensor([out.shape[0]], dtype=torch.int64, device=out.device)
    cnts = [torch.empty_like(cnt) for _ in range(ws)]
    dist.all_gather(cnts, cnt, group=group)
    counts = [int(c.item()) for c in cnts]
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
