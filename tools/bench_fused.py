#!/usr/bin/env python
"""Fused steps either side of the rasterizer against the torch-op chains of the reference they
replace: CUDA-event times of forward + backward, median of 30 after 5 warm-ups.  One JSON line per case."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import ref_torch_ops as ref  # noqa: E402
from bloomscene_b200.fused import l1_ssim_loss, neural_gaussians  # noqa: E402

dev = torch.device("cuda:0")


def timed(fn, iters=30, warm=5):
    ts = []
    for i in range(iters + warm):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        if i >= warm:
            ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


for H, W in [(512, 512), (1080, 1920)]:
    gt = torch.rand(3, H, W, device=dev)
    img = (gt + 0.1 * torch.randn_like(gt)).clamp(0, 1).requires_grad_(True)

    def run(f):
        img.grad = None
        f(img, gt, 0.2).backward()

    t_f, t_r = timed(lambda: run(l1_ssim_loss)), timed(lambda: run(ref.l1_ssim_reference))
    print(json.dumps({"op": "l1_ssim_loss fwd+bwd", "shape": [3, H, W], "fused_ms": round(t_f, 4), "torch_ops_ms": round(t_r, 4),
                      "speedup": round(t_r / t_f, 2)}), flush=True)

for N, K in [(50_000, 10), (200_000, 10)]:
    g = torch.Generator().manual_seed(0)
    ins = [torch.randn(N, 3, generator=g), torch.rand(N, 6, generator=g), torch.randn(N, K, 3, generator=g),
           torch.randn(N * K, 1, generator=g), torch.rand(N * K, 3, generator=g), torch.randn(N * K, 7, generator=g)]
    ins = [t.to(dev).requires_grad_(True) for t in ins]

    def run(f):
        for t in ins:
            t.grad = None
        out = f(*ins)
        sum(o.sum() for o in out[:5]).backward()

    t_f, t_r = timed(lambda: run(neural_gaussians)), timed(lambda: run(ref.neural_gaussians_reference))
    print(json.dumps({"op": "neural_gaussians fwd+bwd (incl. the .sum() losses)", "N": N, "K": K, "fused_ms": round(t_f, 4),
                      "torch_ops_ms": round(t_r, 4), "speedup": round(t_r / t_f, 2)}), flush=True)
