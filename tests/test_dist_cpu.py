"""N > 1 host logic on CPU: world_size-2 gloo processes shard utterances and gather ragged mels to rank 0."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, lens, ret):
    sys.path.insert(0, ROOT)
    from fcl_taco2_b200 import dist as fdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = fdist.my_shard(lens)
        # every utterance's "mel" is a block of rows filled with its global index
        out = torch.cat([torch.full((int(lens[i]), 4), float(i)) for i in mine]) if mine else torch.empty((0, 4))
        bufs = fdist.gather_mels(out, dst=0)
        if rank == 0:
            shards = [fdist.my_shard(lens, r, world) for r in range(world)]
            per_rank = []
            for r in range(world):
                o, outs = 0, []
                for i in shards[r]:
                    outs.append(bufs[r][o:o + int(lens[i])])
                    o += int(lens[i])
                per_rank.append(outs)
            res = fdist.scatter_results(shards, per_rank, len(lens))
            ok = all(res[i].shape == (int(lens[i]), 4) and bool((res[i] == float(i)).all()) for i in range(len(lens)))
            ret.put(("ok" if ok else "mismatch", [len(s) for s in shards]))
        else:
            assert bufs is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_shard_and_gather_world2():
    lens = np.random.RandomState(0).randint(5, 60, size=23)
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lens, ret)) for r in range(2)]
    for p in procs:
        p.start()
    status, counts = ret.get(timeout=100)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert status == "ok" and sum(counts) == 23 and min(counts) >= 10


def _chunk_worker(rank, world, port, lens, ret, shards=None):
    sys.path.insert(0, ROOT)
    from fcl_taco2_b200 import dist as fdist
    from fcl_taco2_b200.plan import output_chunks
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        K = 3
        shard_of = (lambda r: shards[r]) if shards is not None else (lambda r: fdist.my_shard(lens, r, world))
        mine = shard_of(rank)
        ufo = np.concatenate([[0], np.cumsum([int(lens[i]) for i in mine])])
        chunks = output_chunks(ufo, K)
        g = fdist.ChunkedGather(fdist.chunk_bounds(ufo, K), 4, torch.device("cpu"))
        out = torch.zeros((int(ufo[-1]), 4))
        for k, (u0, u1, f0, f1) in enumerate(chunks):      # "compute" chunk k, then hand it to the gather
            for u in range(u0, u1):
                out[int(ufo[u]):int(ufo[u + 1])] = float(mine[u])
            g.on_chunk(k, out, f0, f1)
        bufs = g.finish()
        if rank == 0:
            ok = True
            for r in range(world):
                want = torch.cat([torch.full((int(lens[i]), 4), float(i)) for i in shard_of(r)])
                ok = ok and bufs[r].shape == want.shape and bool((bufs[r] == want).all())
            ret.put("ok" if ok else "mismatch")
        else:
            assert bufs is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_chunked_gather_world2():
    lens = np.random.RandomState(1).randint(5, 60, size=19)
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_chunk_worker, args=(r, 2, port, lens, ret)) for r in range(2)]
    for p in procs:
        p.start()
    status = ret.get(timeout=100)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert status == "ok"


@pytest.mark.timeout(120)
def test_chunked_gather_skewed_shards_world2():
    """The root's shard is ONE long utterance (a single chunk) while the peer has three chunks: the root must still
    receive the peer's chunks 1 and 2 (it used to post receives only for as many chunks as it had itself)."""
    lens = np.array([100, 5, 5, 5, 5, 5, 5])
    shards = [[0], [1, 2, 3, 4, 5, 6]]
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_chunk_worker, args=(r, 2, port, lens, ret, shards)) for r in range(2)]
    for p in procs:
        p.start()
    status = ret.get(timeout=100)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert status == "ok"


def _pipe_worker(rank, world, port, lens, ret):
    sys.path.insert(0, ROOT)
    from fcl_taco2_b200 import dist as fdist
    from fcl_taco2_b200.plan import output_chunks
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        K = 3
        mine = fdist.my_shard(lens)
        ufo = np.concatenate([[0], np.cumsum([int(lens[i]) for i in mine])])
        chunks = output_chunks(ufo, K)
        g = fdist.PipelinedGather(fdist.chunk_bounds(ufo, K), 4, torch.device("cpu"))
        seen = []
        for step in range(5):                              # five passes through two buffer sets, never waiting in between
            out = torch.zeros((int(ufo[-1]), 4))
            cb = g.begin()
            for k, (u0, u1, f0, f1) in enumerate(chunks):
                for u in range(u0, u1):
                    out[int(ufo[u]):int(ufo[u + 1])] = float(mine[u]) + 1000.0 * step
                cb(k, out, f0, f1)
        bufs = g.drain()
        if rank == 0:
            ok = True
            for r in range(world):
                want = torch.cat([torch.full((int(lens[i]), 4), float(i) + 4000.0) for i in fdist.my_shard(lens, r, world)])
                ok = ok and bufs[r].shape == want.shape and bool((bufs[r] == want).all())
            ret.put("ok" if ok else "mismatch")
        else:
            assert bufs is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_pipelined_gather_world2():
    """The gather of pass i is waited for only when its buffer set is reused (pass i+2) or at drain()."""
    lens = np.random.RandomState(3).randint(5, 60, size=17)
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pipe_worker, args=(r, 2, port, lens, ret)) for r in range(2)]
    for p in procs:
        p.start()
    status = ret.get(timeout=100)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert status == "ok"


def test_output_chunks_cover_everything():
    sys.path.insert(0, ROOT)
    from fcl_taco2_b200.plan import output_chunks
    rs = np.random.RandomState(2)
    for B, K in ((1, 4), (3, 4), (50, 4), (50, 1), (7, 16)):
        ufo = np.concatenate([[0], np.cumsum(rs.randint(1, 90, size=B))])
        ch = output_chunks(ufo, K)
        assert 1 <= len(ch) <= K and ch[0][0] == 0 and ch[0][2] == 0 and ch[-1][1] == B and ch[-1][3] == ufo[-1]
        for a, b in zip(ch[:-1], ch[1:]):
            assert a[1] == b[0] and a[3] == b[2]
        for u0, u1, f0, f1 in ch:
            assert f0 == ufo[u0] and f1 == ufo[u1] and f1 > f0
