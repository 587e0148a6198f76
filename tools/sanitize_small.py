#!/usr/bin/env python
"""Small forward+backward scenes for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import parity_lib as pl
from bloomscene_b200 import synthetic
from bloomscene_b200.multiview import GaussianParams, view_sharded_step

dev = torch.device("cuda:0")
api = pl.ours()
for (P, color, W, H, mu) in [(3000, "sh3", 160, 96, -3.2), (1500, "precomp", 70, 50, -2.6), (40, "sh0", 33, 17, -2.0)]:
    scene = synthetic.make_scene(P, "object", color, mu, seed=3).to(dev)
    cam = synthetic.orbit_camera(W, H, 0.4).to(dev)
    Wc, Wd = [t.to(dev) for t in synthetic.loss_weights(W, H)]
    bg = torch.tensor([0.1, 0.2, 0.3], device=dev)
    out = pl.run_autograd(api, scene, cam, bg, Wc, Wd)
    torch.cuda.synchronize()
    print(P, color, float(out["color"].sum()), {k: float(v.abs().sum()) for k, v in out["grads"].items() if v is not None})
    # multi-stream step with gradient sinks
    params = GaussianParams(scene)
    cams = [synthetic.orbit_camera(W, H, y).to(dev) for y in (0.0, 0.9, 1.7, 2.9)]
    r = view_sharded_step(params, cams, bg, api.GaussianRasterizer, lambda c, d, vi: (c * Wc).sum() + (d * Wd).sum())
    torch.cuda.synchronize()
    print("step loss", float(r["loss"]))
print("done")
