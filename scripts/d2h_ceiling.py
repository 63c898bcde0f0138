#!/usr/bin/env python
"""Host-side ceiling of the end-to-end loop at N GPUs: aggregate pinned device->host bandwidth when every rank streams the per-step
read-back (13.1 MB at 4096 envs) at once, against rank 0 streaming alone.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 scripts/d2h_ceiling.py
"""
import json
import os

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
res = {}
for mb in (0.8, 13.1, 64.0):
    n = int(mb * 1e6)
    src = torch.empty(n, dtype=torch.uint8, device="cuda")
    dst = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(2)]
    for phase in ("alone", "all"):
        active = rank == 0 or phase == "all"
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 200
        e0.record()
        if active:
            for i in range(reps):
                dst[i & 1].copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        nact = 1 if phase == "alone" else world
        res[f"{mb}MB_{phase}"] = {"GB/s_aggregate": round(nact * reps * n / (ms.item() * 1e-3) / 1e9, 1), "ms_per_copy": round(ms.item() / reps, 4)}
if rank == 0:
    print(json.dumps({"world": world, "d2h": res}))
if world > 1:
    dist.destroy_process_group()
