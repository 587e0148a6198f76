#!/usr/bin/env python
"""Where a stacked forward (brs_forward_views) spends its time: python tools/stack_probe.py B 2,8"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import parity_lib as pl
from workload import synthetic

name = sys.argv[1]
stacks = [int(x) for x in sys.argv[2].split(",")]
dev = torch.device("cuda:0")
api = pl.ours(); _C = api._C
cfg = synthetic.CONFIGS[name]
scene = synthetic.config_scene(name).to(dev)
cams = [c.to(dev) for c in synthetic.config_cameras(name, 64)]
bg = torch.zeros(3, device=dev)
e = torch.Tensor([])
opt = lambda t: e if t is None else t
for n in stacks:
    groups = [cams[j:j + n] for j in range(0, len(cams) - n + 1, n)]
    def call(grp):
        vm = torch.stack([v.viewmatrix for v in grp]); pm = torch.stack([v.projmatrix for v in grp]); cp = torch.stack([v.campos for v in grp])
        return _C.rasterize_gaussians_views(bg, scene.means3D, opt(scene.colors_precomp), scene.opacities, scene.scales, scene.rotations,
                                            1.0, e, vm, pm, [grp[0].tanfovx], [grp[0].tanfovy], cfg["H"], cfg["W"], opt(scene.shs),
                                            scene.sh_degree, cp, False, False)
    for g in groups[:2]: call(g)
    torch.cuda.synchronize(); _C.forward_stats(True)
    _C.stage_times(); _C.stage_timing(True)
    for g in groups: call(g)
    t = _C.stage_times(); _C.stage_timing(False)
    print(name, "stack", n, "stage ms per stack", {k: round(v[0] / len(groups), 4) for k, v in t.items() if v[0] > 0},
          "sum", round(sum(v[0] for v in t.values()) / len(groups), 4))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for g in groups: call(g)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("   wall ms per stack", round(dt / len(groups) * 1e3, 4), "per view", round(dt / len(groups) / n * 1e3, 4), "stats", _C.forward_stats(True))
    t0 = time.perf_counter()
    for g in groups:
        call(g); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("   wall ms per stack, sync after each", round(dt / len(groups) * 1e3, 4))
