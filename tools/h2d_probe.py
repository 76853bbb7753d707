"""Host-link probe (run under torchrun, one rank per GPU): pinned-host <-> device copy bandwidth of every rank ALONE and of
all ranks TOGETHER. Answers whether the end-to-end rate at N GPUs (bench.py `e2e`) is limited by the box's host links or by
the code: e2e moves 2 x 5.56 GB per GPU and step; nothing in the library touches those bytes except cudaMemcpyAsync.
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/h2d_probe.py"""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 29                                   # 4 GiB of float64
h = torch.empty(n, dtype=torch.float64).pin_memory()
d = torch.empty(n, dtype=torch.float64, device="cuda")
h.fill_(1.0)


def rate(fn, reps=3):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return reps * n * 8 / (time.perf_counter() - t0) / 1e9


def both():
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)


h2 = torch.empty(n // 2, dtype=torch.float64).pin_memory()
d2 = torch.empty(n // 2, dtype=torch.float64, device="cuda")
out = {}
# alone: one rank at a time
alone = []
for r in range(world):
    if r == rank:
        v = rate(lambda: d.copy_(h, non_blocking=True))
    else:
        v = 0.0
        torch.cuda.synchronize()
    if world > 1:
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        v = float(t.item())
    alone.append(v)
tog_h2d = rate(lambda: d.copy_(h, non_blocking=True))
tog_d2h = rate(lambda: h.copy_(d, non_blocking=True))
vals = torch.tensor([tog_h2d, tog_d2h], device="cuda", dtype=torch.float64)
if world > 1:
    dist.all_reduce(vals)
if rank == 0:
    print(json.dumps({"gpus": world, "h2d_alone_GBs_per_rank": alone, "h2d_together_aggregate_GBs": float(vals[0]),
                      "d2h_together_aggregate_GBs": float(vals[1]), "cpus": os.cpu_count()}), flush=True)
if world > 1:
    dist.destroy_process_group()
