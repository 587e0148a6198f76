#!/usr/bin/env python
"""Where does a view-sharded step spend GPU time?  torch.profiler kernel table of one 8-view step
(config E views), to separate the native kernels from the torch glue around them (loss, autograd
accumulation, fills).  Diagnostic only; not a bench.

    python tools/step_trace.py [--views 8]
"""
from __future__ import annotations

import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--views", type=int, default=8)
    a = ap.parse_args()
    import bloomscene_b200
    from workload import synthetic
    from bloomscene_b200.multiview import GaussianParams, view_sharded_step

    api = bloomscene_b200._api
    dev = torch.device("cuda:0")
    cfg = synthetic.CONFIGS["E"]
    scene = synthetic.config_scene("E").to(dev)
    cams = [c.to(dev) for c in synthetic.config_cameras("E", a.views)]
    Wc, Wd = [t.to(dev) for t in synthetic.loss_weights(cfg["W"], cfg["H"])]
    bg = torch.zeros(3, device=dev)
    params = GaussianParams(scene)
    # loss = <color, Wc> + <depth, Wd>, written as two dot products: one reduction kernel
    # forward and one scaling kernel backward per term, for both arms alike
    Wc_flat, Wd_flat = Wc.reshape(-1), Wd.reshape(-1)
    loss_fn = lambda color, depth, vi: torch.dot(color.reshape(-1), Wc_flat) + torch.dot(depth.reshape(-1), Wd_flat)
    step = lambda: view_sharded_step(params, cams, bg, api.GaussianRasterizer, loss_fn)
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step()
    e1.record()
    torch.cuda.synchronize()
    print(f"step of {a.views} views: {e0.elapsed_time(e1) / a.views:.3f} ms/view (untraced)")
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    if os.environ.get("STEP_TRACE_JSON"):
        prof.export_chrome_trace(os.environ["STEP_TRACE_JSON"])
    rows = []
    for ev in prof.key_averages():
        t = getattr(ev, "device_time_total", None)
        if t is None:
            t = getattr(ev, "cuda_time_total", 0)
        if t and ev.device_type.name != "CPU":
            rows.append((t, ev.count, ev.key))
    rows.sort(reverse=True)
    tot = sum(r[0] for r in rows)
    print(f"GPU kernel time per view: {tot / a.views / 1e3:.3f} ms")
    for t, n, k in rows[:40]:
        print(f"{t / a.views:9.1f} us/view  x{n / a.views:5.1f}  {k[:110]}")
    # host side: how long the CPU spends per view
    cpu = [(ev.self_cpu_time_total, ev.count, ev.key) for ev in prof.key_averages() if ev.self_cpu_time_total > 0]
    cpu.sort(reverse=True)
    print("host self time per view (top 12):")
    for t, n, k in cpu[:12]:
        print(f"{t / a.views:9.1f} us/view  x{n / a.views:5.1f}  {k[:100]}")


if __name__ == "__main__":
    main()
