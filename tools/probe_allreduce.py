#!/usr/bin/env python
"""Time the step's one collective (sum all-reduce of a 236 MB fp32 bucket) on its own, under whatever NCCL_* settings the
environment carries: torchrun --nproc-per-node N tools/probe_allreduce.py"""
import json
import os

import torch
import torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
world = dist.get_world_size()
buf = torch.zeros(59_000_002, device="cuda")
for _ in range(5):
    dist.all_reduce(buf)
torch.cuda.synchronize()
dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    dist.all_reduce(buf)
b.record()
torch.cuda.synchronize()
ms = torch.tensor([a.elapsed_time(b) / 20], device="cuda")
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if dist.get_rank() == 0:
    n = buf.numel() * 4
    print(json.dumps({"world": world, "bytes": n, "ms": round(ms.item(), 4), "busbw_GBps": round(2 * (world - 1) / world * n / (ms.item() * 1e-3) / 1e9, 1),
                      "env": {k: v for k, v in os.environ.items() if k.startswith("NCCL_")}}))
dist.destroy_process_group()
