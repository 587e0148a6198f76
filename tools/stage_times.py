#!/usr/bin/env python
"""Per-stage CUDA-event times of the forward (and backward) for one config: python tools/stage_times.py B [--bwd]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import parity_lib as pl
from workload import synthetic

name = sys.argv[1]
bwd = "--bwd" in sys.argv
dev = torch.device("cuda:0")
api = pl.ours(); _C = api._C
cfg = synthetic.CONFIGS[name]
scene = synthetic.config_scene(name).to(dev)
cams = [c.to(dev) for c in synthetic.config_cameras(name, 8)]
bg = torch.zeros(3, device=dev)
Wc = synthetic.loss_weights(cfg["W"], cfg["H"])[0].to(dev)
e = torch.Tensor([])
def one(cam):
    out = _C.rasterize_gaussians(*pl.forward_args(scene, cam, bg))
    if bwd:
        R, color, depth, radii, geom, binning, img = out
        _C.rasterize_gaussians_backward(bg, scene.means3D, radii, scene.colors_precomp if scene.colors_precomp is not None else e,
            scene.scales, scene.rotations, 1.0, e, cam.viewmatrix, cam.projmatrix, cam.tanfovx, cam.tanfovy, Wc, e,
            scene.shs if scene.shs is not None else e, scene.sh_degree, cam.campos, geom, R, binning, img, False)
    return out
for c in cams[:2]: one(c)
torch.cuda.synchronize(); _C.stage_times(); _C.stage_timing(True)
for c in cams: one(c)
t = _C.stage_times(); _C.stage_timing(False)
print(name, {k: round(v[0] / len(cams), 4) for k, v in t.items()}, "sum", round(sum(v[0] for v in t.values()) / len(cams), 4))
torch.cuda.synchronize(); t0 = time.perf_counter()
for c in cams: one(c)
torch.cuda.synchronize(); print("wall ms/view", round((time.perf_counter() - t0) / len(cams) * 1e3, 4))
