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


class ChunkedGather:
    """Final gather of the ragged mels to `dst`, overlapped with the tail of the pass: the postnet is launched per
    group of utterances (Engine.run_uploaded(out_chunks=K, chunk_cb=gather.on_chunk)) and each finished group is sent
    while the next one still computes. The frame ranges of every rank's groups are exchanged up front (K+1 integers per
    rank; with forced durations they are known on the host before the pass). The root's NVLink ingress
    ((N-1) x F x odim x 4 bytes) stays the floor; what is hidden is the postnet time."""

    def __init__(self, bounds, odim, device, dst: int = 0, group=None, dtype=torch.float32):
        """bounds: this rank's chunk frame boundaries [f_0=0, f_1, ..., f_K=F] (same K on every rank)."""
        self.ws, self.rank, self.dst, self.group = dist.get_world_size(group), dist.get_rank(group), dst, group
        mine = torch.tensor(list(bounds), dtype=torch.int64, device=device)
        allb = [torch.empty_like(mine) for _ in range(self.ws)]
        dist.all_gather(allb, mine, group=group)
        self.bounds = [b.cpu().tolist() for b in allb]
        self.reqs = []
        self.bufs = None
        if self.rank == dst:
            self.bufs = [None if r == dst else torch.empty((self.bounds[r][-1], odim), dtype=dtype, device=device)
                         for r in range(self.ws)]

    def on_chunk(self, k, out, f_lo, f_hi):
        """Called by the engine right after chunk k's kernels were enqueued on the current stream."""
        if self.ws == 1:
            return
        if self.rank == self.dst:
            self.bufs[self.dst] = out
            ops = []
            for r in range(self.ws):
                lo, hi = self.bounds[r][k], self.bounds[r][k + 1]
                if r != self.dst and hi > lo:
                    ops.append(dist.P2POp(dist.irecv, self.bufs[r][lo:hi], r, self.group))
        else:
            ops = [dist.P2POp(dist.isend, out[f_lo:f_hi], self.dst, self.group)] if f_hi > f_lo else []
        if ops:
            self.reqs.extend(dist.batch_isend_irecv(ops))

    def finish(self):
        """Make the current stream wait for all transfers. -> per-rank tensors on dst, else None."""
        for r in self.reqs:
            r.wait()
        self.reqs = []
        return self.bufs
