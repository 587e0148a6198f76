#!/usr/bin/env python
"""Timeline of ONE view-sharded step on rank 0 of an N-rank job (torch profiler / CUPTI): when each lane stream is
busy, when the bucket is zeroed, when the NCCL all-reduce starts and ends, what the step's span is.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
        tools/step_timeline.py --streams 4 > gpurun_out/timeline_n8.md
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--views", type=int, default=64)
    ap.add_argument("--streams", type=int, default=4)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import bloomscene_b200
    from bloomscene_b200.multiview import view_sharded_step
    from workload import synthetic
    from workload.params import GaussianParams

    api = bloomscene_b200._api
    cfg = synthetic.CONFIGS["E"]
    params = GaussianParams(synthetic.config_scene("E").to(dev))
    cams = [c.to(dev) for c in synthetic.config_cameras("E", a.views)]
    Wc, Wd = [t.to(dev).reshape(-1) for t in synthetic.loss_weights(cfg["W"], cfg["H"])]
    bg = torch.zeros(3, device=dev)
    loss_fn = lambda color, depth, vi: torch.dot(color.reshape(-1), Wc) + torch.dot(depth.reshape(-1), Wd)
    step = lambda: view_sharded_step(params, cams, bg, api.GaussianRasterizer, loss_fn, rank=rank, world=world, streams=a.streams)
    for _ in range(4):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    path = os.path.join(tempfile.gettempdir(), f"step_trace_{rank}.json")
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    if rank == 0:
        prof.export_chrome_trace(path)
        ev = [e for e in json.load(open(path))["traceEvents"] if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy")]
        t0 = min(e["ts"] for e in ev)
        end = max(e["ts"] + e["dur"] for e in ev)
        print(f"# One step on rank 0 of {world} ({a.views // world} views, {a.streams} lane streams): GPU span {(end - t0) / 1e3:.3f} ms\n")
        streams = {}
        for e in ev:
            streams.setdefault(e["args"].get("stream"), []).append(e)
        print("| stream | kernels | first start ms | last end ms | busy ms | what |")
        print("|---|---|---|---|---|---|")
        for sid, es in sorted(streams.items(), key=lambda kv: min(x["ts"] for x in kv[1])):
            names = sorted({x["name"].split("(")[0].split("<")[0].split("::")[-1][:28] for x in es})
            busy = sum(x["dur"] for x in es)
            print(f"| {sid} | {len(es)} | {(min(x['ts'] for x in es) - t0) / 1e3:.3f} | {(max(x['ts'] + x['dur'] for x in es) - t0) / 1e3:.3f} | "
                  f"{busy / 1e3:.3f} | {', '.join(names[:6])}{' ...' if len(names) > 6 else ''} |")
        print()
        for key, label in (("nccl", "NCCL all-reduce kernel"), ("Memset", "memsets >= 1 MB")):
            for e in ev:
                if key.lower() in e["name"].lower() and (key == "nccl" or e["args"].get("bytes", 0) >= 1 << 20):
                    print(f"* {label}: start {(e['ts'] - t0) / 1e3:.3f} ms, duration {e['dur'] / 1e3:.3f} ms ({e['name'][:60]})")
        last_pre = max((e["ts"] + e["dur"] for e in ev if "preprocess_backward" in e["name"]), default=t0)
        print(f"* last preprocess_backward ends at {(last_pre - t0) / 1e3:.3f} ms; the step ends at {(end - t0) / 1e3:.3f} ms")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
