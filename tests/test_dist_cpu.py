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
