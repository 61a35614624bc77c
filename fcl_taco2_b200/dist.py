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


def chunk_bounds(utt_frame_off, k: int):
    """Frame boundaries [0, f_1, ..., F] of the <= k output chunks of a shard, padded to exactly k + 1 entries
    (trailing empty chunks) so that every rank hands ChunkedGather the same number of bounds."""
    from .plan import output_chunks
    ch = output_chunks(utt_frame_off, k)
    bounds = ([c[2] for c in ch] + [ch[-1][3]]) if ch else [0]
    return bounds + [bounds[-1]] * (k + 1 - len(bounds))


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
        self.next_k = 0
        self.n_chunks = len(self.bounds[0]) - 1
        if any(len(b) != self.n_chunks + 1 for b in self.bounds):
            raise ValueError("ChunkedGather: every rank must pass the same number of chunk bounds")
        if self.rank == dst:
            self.bufs = [None if r == dst else torch.empty((self.bounds[r][-1], odim), dtype=dtype, device=device)
                         for r in range(self.ws)]

    def _post(self, k, out):
        """Post the transfers of chunk k: the root receives every peer's chunk k (its extent comes from the exchanged
        bounds, NOT from the root's own chunk list), a peer sends its own."""
        ops = []
        if self.rank == self.dst:
            for r in range(self.ws):
                lo, hi = self.bounds[r][k], self.bounds[r][k + 1]
                if r != self.dst and hi > lo:
                    ops.append(dist.P2POp(dist.irecv, self.bufs[r][lo:hi], r, self.group))
        else:
            lo, hi = self.bounds[self.rank][k], self.bounds[self.rank][k + 1]
            if hi > lo:
                ops.append(dist.P2POp(dist.isend, out[lo:hi], self.dst, self.group))
        if ops:
            self.reqs.extend(dist.batch_isend_irecv(ops))

    def on_chunk(self, k, out, f_lo, f_hi):
        """Called by the engine right after chunk k's kernels were enqueued on the current stream."""
        if self.ws == 1:
            return
        if k != self.next_k:
            raise RuntimeError(f"ChunkedGather: chunk {k} handed over out of order (expected {self.next_k})")
        if (f_lo, f_hi) != (self.bounds[self.rank][k], self.bounds[self.rank][k + 1]):
            raise RuntimeError(f"ChunkedGather: chunk {k} covers frames [{f_lo}, {f_hi}) but the bounds exchanged at "
                               f"construction say [{self.bounds[self.rank][k]}, {self.bounds[self.rank][k + 1]})")
        if self.rank == self.dst:
            self.bufs[self.dst] = out
        self._post(k, out)
        self.next_k = k + 1

    def finish(self):
        """Make the current stream wait for all transfers. -> per-rank tensors on dst, else None.
        Ranks may have DIFFERENT numbers of non-empty chunks (ragged or tiny shards): the root posts the receives of
        every chunk its own pass did not reach, so a peer's later chunks are never left unmatched."""
        if self.ws > 1:
            for k in range(self.next_k, self.n_chunks):
                self._post(k, None)            # a rank's own chunks beyond next_k are empty (bounds are padded)
        for r in self.reqs:
            r.wait()
        self.reqs = []
        self.next_k = 0                        # ready for the next pass over the same bounds
        return self.bufs



class PipelinedGather:
    """Takes the final gather off the critical path of a STREAM of passes: the mels of pass i travel to the root while
    pass i+1 computes (two receive-buffer sets, alternated), exactly as `Tacotron2_sa.inference_stream` overlaps the
    device->host copy of a batch with the next batch. Only the last pass's transfer is exposed.

    Use: `g = PipelinedGather(bounds, odim, device)`; per pass `cb = g.begin()` is handed to
    `Engine.run_uploaded(..., out_chunks=K, chunk_cb=cb)`; `g.drain()` after the last pass waits for everything and
    returns the most recent per-rank buffers on the root (None elsewhere). All passes must have the same bounds
    (a benchmark loop, or a decode driver that exchanges bounds per batch and builds one object per shape)."""

    def __init__(self, bounds, odim, device, dst: int = 0, group=None, dtype=torch.float32, depth: int = 2):
        self.sets = [ChunkedGather(bounds, odim, device, dst, group, dtype) for _ in range(depth)]
        self.n, self.last = 0, None
        self.keep = [None] * depth            # the output tensor of a pass stays alive until its transfer was waited for
        self.busy = [False] * depth           # a pass was started on this set and not yet finished

    def begin(self):
        slot = self.n % len(self.sets)
        g = self.sets[slot]
        if self.busy[slot]:
            g.finish()                        # the transfer that used this buffer set `depth` passes ago (long done)
        self.busy[slot] = True
        self.n += 1
        self.last = g

        def cb(k, out, f_lo, f_hi):
            self.keep[slot] = out
            g.on_chunk(k, out, f_lo, f_hi)
        return cb

    kind = "nccl-p2p (ncclSend/ncclRecv), pipelined across steps"

    def close(self):
        self.drain()

    def drain(self):
        bufs = None
        for i in range(len(self.sets)):       # oldest first
            slot = (self.n + i) % len(self.sets)
            g = self.sets[slot]
            if self.busy[slot]:
                b = g.finish()
                self.busy[slot] = False
                if g is self.last:
                    bufs = b
        self.keep = [None] * len(self.sets)
        return bufs


class _DevMem:
    """Zero-copy torch view of a raw device pointer (peer-mapped or cudaMalloc'ed by the C library)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class PeerDmaGather:
    """Same contract as PipelinedGather (begin() -> chunk callback, drain()), but the mels travel by copy-engine DMA into
    receive buffers the root exports through CUDA IPC (csrc/peer.cu): every rank maps the root's buffers once, then per
    pass pushes each finished postnet chunk with `fcl_copy_async` on a side stream -- over NVLink, with no SM kernel on
    either side, so the persistent decoder of the NEXT pass keeps every SM while the previous pass's mels move.
    Hand-shake per buffer set s (two sets, alternated per pass), all in root memory:
        free[s]   root -> peers  "set s may be overwritten by pass n"   (root writes n+1 at begin() of pass n)
        done[s,r] peer r -> root "my mels of pass n are in set s"        (written after the peer's last copy)
    `drain()` / the reuse of a set make the root's stream wait for done == n+1 (one-warp polling kernel)."""

    kind = "peer-dma (CUDA IPC receive buffers on the root, copy-engine pushes over NVLink, pipelined across steps)"

    def __init__(self, bounds, odim, device, dst: int = 0, group=None, depth: int = 2):
        import ctypes as C
        from . import _lib
        self.lib, self.C = _lib, C
        self.ws, self.rank, self.dst, self.group = dist.get_world_size(group), dist.get_rank(group), dst, group
        self.device, self.odim, self.depth = torch.device(device), odim, depth
        mine = torch.tensor(list(bounds), dtype=torch.int64, device=device)
        allb = [torch.empty_like(mine) for _ in range(self.ws)]
        dist.all_gather(allb, mine, group=group)
        self.bounds = [b.cpu().tolist() for b in allb]
        self.n_chunks = len(self.bounds[0]) - 1
        self.row_bytes = odim * 4
        self.n = 0
        self.keep = [None] * depth
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.owned, self.mapped = [], []
        # layout of the root's export, per set: [flags: free(1) + done(ws) int32, padded to 256 B][rank 0 .. ws-1 mel buffers]
        self.flag_bytes = 256 * ((4 * (1 + self.ws) + 255) // 256)
        offs, o = [], self.flag_bytes
        for r in range(self.ws):
            offs.append(o)
            o += 256 * ((self.bounds[r][-1] * self.row_bytes + 255) // 256) if r != dst else 0
        self.offs, self.set_bytes = offs, max(o, self.flag_bytes)
        with torch.cuda.device(self.device):
            handles = [None] * depth
            if self.rank == dst:
                self.base = []
                for s in range(depth):
                    p = C.c_void_p()
                    _lib.plain("fcl_peer_alloc", C.c_int64(self.set_bytes), C.byref(p))
                    self.owned.append(p.value)
                    self.base.append(p.value)
                    h = C.create_string_buffer(64)
                    _lib.plain("fcl_ipc_export", C.c_void_p(p.value), h)
                    handles[s] = h.raw
            obj = [handles]
            dist.broadcast_object_list(obj, src=dst, group=group)
            if self.rank != dst:
                self.base = []
                for s in range(depth):
                    p = C.c_void_p()
                    _lib.plain("fcl_ipc_open", obj[0][s], C.byref(p))
                    self.mapped.append(p.value)
                    self.base.append(p.value)
        self.bufs = None
        if self.rank == dst:
            self.views = [[None if r == dst else
                           torch.as_tensor(_DevMem(self.base[s] + offs[r], (self.bounds[r][-1], odim), "<f4"), device=self.device)
                           for r in range(self.ws)] for s in range(depth)]
        self.pending = [0] * depth            # sequence number (n+1) of the pass in flight on each set, 0 = idle

    def _stream(self, s):
        return self.C.c_void_p(s.cuda_stream)

    def _wait_done(self, s):
        """Root: the current stream waits until every peer's pass on set s has landed."""
        if self.pending[s] and self.rank == self.dst and self.ws > 1:
            peers = [r for r in range(self.ws) if r != self.dst and self.bounds[r][-1] > 0]
            if peers:
                # done flags sit at int32 index 1 + r; wait on the contiguous range (idle peers are written too: see begin)
                self.lib.plain("fcl_wait_flags", self.C.c_void_p(self.base[s] + 4), self.C.c_int32(self.ws),
                               self.C.c_int32(self.pending[s]), self._stream(torch.cuda.current_stream(self.device)))
        self.pending[s] = 0

    def begin(self):
        s, seq = self.n % self.depth, self.n + 1
        self.n += 1
        main = torch.cuda.current_stream(self.device)
        last_k = [None]
        if self.rank == self.dst:
            self._wait_done(s)                                        # pass n - depth fully received (and, by contract, consumed)
            # root's own done flag + "free" in one store sequence: free[s] = seq, done[s, dst] = seq
            self.lib.plain("fcl_write_flags", self.C.c_void_p(self.base[s]), self.C.c_int32(1), self.C.c_int32(seq), self._stream(main))
            self.lib.plain("fcl_write_flags", self.C.c_void_p(self.base[s] + 4 * (1 + self.dst)), self.C.c_int32(1), self.C.c_int32(seq),
                           self._stream(main))
            self.pending[s] = seq
            self.cur = s

            def cb(k, out, f_lo, f_hi):
                self.keep[s] = out
            return cb
        cs = self.copy_stream
        # the copies of this pass may overwrite set s only after the root released it
        self.lib.plain("fcl_wait_flags", self.C.c_void_p(self.base[s]), self.C.c_int32(1), self.C.c_int32(seq), self._stream(cs))
        dst_base = self.base[s] + self.offs[self.rank]

        def cb(k, out, f_lo, f_hi):
            self.keep[s] = out
            out.record_stream(cs)                   # the allocator must not recycle `out` while a queued copy still reads it
            if (f_lo, f_hi) != (self.bounds[self.rank][k], self.bounds[self.rank][k + 1]):
                raise RuntimeError(f"PeerDmaGather: chunk {k} covers frames [{f_lo}, {f_hi}) but the exchanged bounds differ")
            ev = torch.cuda.Event()
            ev.record(main)
            cs.wait_event(ev)
            self.lib.plain("fcl_copy_async", self.C.c_void_p(dst_base + f_lo * self.row_bytes),
                           self.C.c_void_p(out.data_ptr() + f_lo * self.row_bytes), self.C.c_int64((f_hi - f_lo) * self.row_bytes),
                           self._stream(cs))
            if f_hi == self.bounds[self.rank][-1]:                    # last non-empty chunk: publish "done"
                self.lib.plain("fcl_write_flags", self.C.c_void_p(self.base[s] + 4 * (1 + self.rank)), self.C.c_int32(1),
                               self.C.c_int32(seq), self._stream(cs))
        if self.bounds[self.rank][-1] == 0:                           # an empty shard still reports "done"
            self.lib.plain("fcl_write_flags", self.C.c_void_p(self.base[s] + 4 * (1 + self.rank)), self.C.c_int32(1),
                           self.C.c_int32(seq), self._stream(cs))
        return cb

    def drain(self):
        """Wait (stream-wise on the root, for the local copies elsewhere) for every pass in flight. -> per-rank buffers of
        the most recent pass on the root (entry `dst` = the root's own output tensor), None elsewhere."""
        if self.rank == self.dst:
            for i in range(self.depth):
                self._wait_done((self.n + i) % self.depth)
            if self.n == 0:
                return None
            s = (self.n - 1) % self.depth
            out = list(self.views[s])
            out[self.dst] = self.keep[s]
            return out
        torch.cuda.current_stream(self.device).wait_stream(self.copy_stream)
        return None

    def close(self):
        torch.cuda.synchronize(self.device)
        if self.ws > 1:
            dist.barrier(group=self.group)          # nobody unmaps / frees while a peer may still be writing
        with torch.cuda.device(self.device):
            for p in self.mapped:
                self.lib.plain("fcl_ipc_close", self.C.c_void_p(p))
            if self.ws > 1:
                dist.barrier(group=self.group)
            for p in self.owned:
                self.lib.plain("fcl_peer_free", self.C.c_void_p(p))
        self.mapped, self.owned = [], []


def make_gather(bounds, odim, device, kind: str = "auto", dst: int = 0, group=None):
    """Final-gather transport for a stream of passes: 'peer' = PeerDmaGather, 'nccl' = PipelinedGather over
    ncclSend/ncclRecv, 'auto' = peer when every rank can set it up (decided collectively), else nccl."""
    g = None
    if kind in ("auto", "peer"):
        ok = 1
        try:
            g = PeerDmaGather(bounds, odim, device, dst, group)
        except Exception as e:                       # no IPC / no peer access on this box
            if kind == "peer":
                raise
            ok, g = 0, None
            import sys
            print(f"[fcl dist] peer-DMA gather unavailable on rank {dist.get_rank(group)}: {e!r}", file=sys.stderr)
        flag = torch.tensor([ok], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 1:
            return g
        g = None
    return PipelinedGather(bounds, odim, device, dst, group)


def bind_to_gpu_numa_node(local_rank: int):
    """Pin this process to the CPUs NVML reports as local to its GPU, BEFORE any pinned host buffer is allocated:
    first-touch then places the staging buffers on the GPU's own NUMA node (8 ranks x 185 MB of D2H per pass otherwise
    funnel through one socket's memory controllers). Best effort: -> the previous affinity set (to restore with
    os.sched_setaffinity) when the binding was applied, None when NVML / affinity is unavailable or nothing changed."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        idx = local_rank
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if local_rank < len(ids) and ids[local_rank].isdigit():
                idx = int(ids[local_rank])
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * w + b for w, mask in enumerate(words) for b in range(64) if (int(mask) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            return None
        os.sched_setaffinity(0, cpus)
        return allowed
    except Exception:
        return None
