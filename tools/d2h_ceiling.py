#!/usr/bin/env python
"""memcpy-only ceiling of the end-to-end path's device->host leg: N ranks each copy one batch of mels
(185 MB fp32, pinned destination) per step, nothing else running. Launch under torchrun with 8 ranks; prints one JSON
line with the aggregate GB/s at 1, 2, 4, 8 active ranks. (is e2e at N = 8 bound by the host?)"""
import json, os, sys, time
import torch, torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
nbytes = 579359 * 80 * 4
src = torch.empty(nbytes, dtype=torch.uint8, device=dev).fill_(rank + 1)
dst = [torch.empty(nbytes, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
out = {}
K = 12
n = 1
while n <= world:
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = 0.0
    if rank < n:
        for w in range(2):
            dst[w].copy_(src, non_blocking=True)
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if rank < n:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(K):
            dst[k & 1].copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out[str(n)] = {"ms_per_copy": float(t[0]) / K, "aggregate_gbs": n * nbytes * K / (float(t[0]) * 1e-3) / 1e9,
                   "per_gpu_gbs": nbytes * K / (float(t[0]) * 1e-3) / 1e9}
    n *= 2
if rank == 0:
    print(json.dumps({"what": "D2H memcpy-only ceiling, 185 MB per rank per copy, pinned destination", "bytes": nbytes,
                      "cpus": os.cpu_count(), "by_active_ranks": out}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
