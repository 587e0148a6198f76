#!/usr/bin/env python
"""Small forward+backward scenes for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import parity_lib as pl
from workload import synthetic
from bloomscene_b200.multiview import view_sharded_step
from workload.params import GaussianParams

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
    # opt-in depth gradient (templated blend backward) and the forward-only batch entry point
    settings = [synthetic.raster_settings(c, scene.sh_degree, bg, api.GaussianRasterizationSettings) for c in cams]
    leaves = {n: getattr(scene, n).detach().clone().requires_grad_(True) for n in ("means3D", "opacities", "scales", "rotations")}
    col = {"shs": scene.shs} if scene.shs is not None else {"colors_precomp": scene.colors_precomp}
    rast = api.GaussianRasterizer(settings[0], depth_gradient=True)
    c_img, _, d_img = rast(means3D=leaves["means3D"], means2D=torch.zeros_like(scene.means3D, requires_grad=True),
                           opacities=leaves["opacities"], scales=leaves["scales"], rotations=leaves["rotations"], **col)
    ((c_img * Wc).sum() + (d_img * Wd).sum()).backward()
    frames = api.render_views(settings, scene.means3D, scene.opacities, scales=scene.scales, rotations=scene.rotations, **col)
    torch.cuda.synchronize()
    # the same frames as ONE stacked pipeline (brs_forward_views), twice: EXACT, then optimistic; then on two lanes
    for lanes in (1, 1, 2):
        stacked = api.render_views(settings, scene.means3D, scene.opacities, scales=scene.scales, rotations=scene.rotations,
                                   stack=3, streams=lanes, **col)
        torch.cuda.synchronize()
        assert torch.equal(stacked[0], frames[0]) and torch.equal(stacked[1], frames[1])
    print("depth-grad", float(leaves["means3D"].grad.abs().sum()), "frames", float(frames[0].sum()))
    # forward modes: optimistic with stale marks (overflow -> re-run), deferred with far too small capacities
    _C = api._C
    e = torch.Tensor([])
    args = pl.forward_args(scene, cam, bg)
    _C.reset_marks()
    _C.rasterize_gaussians_ex(*pl.forward_args(scene, cam, bg, scale_modifier=0.3), _C.FWD_AUTO)
    _C.rasterize_gaussians_ex(*pl.forward_args(scene, cam, bg, scale_modifier=2.5), _C.FWD_AUTO)  # overflows the marks
    report = torch.zeros(8, dtype=torch.int32).pin_memory()
    _C.rasterize_gaussians_ex(*args, _C.FWD_DEFERRED, 16, 8, 8, report)
    torch.cuda.synchronize()
    print("modes", _C.forward_stats(True), "deferred overflow word", int(report[5]))
    _C.reset_marks()

# the steps either side of the rasterizer
from bloomscene_b200.fused import l1_ssim_loss, neural_gaussians
g = torch.Generator().manual_seed(0)
img = torch.rand(3, 45, 37, generator=g).to(dev).requires_grad_(True)
l1_ssim_loss(img, torch.rand(3, 45, 37, generator=g).to(dev)).backward()
N, K = 700, 10
ins = [torch.randn(N, 3, generator=g), torch.rand(N, 6, generator=g), torch.randn(N, K, 3, generator=g),
       torch.randn(N * K, 1, generator=g), torch.rand(N * K, 3, generator=g), torch.randn(N * K, 7, generator=g)]
ins = [t.to(dev).requires_grad_(True) for t in ins]
out = neural_gaussians(*ins)
sum(o.sum() for o in out[:5]).backward()
rs = synthetic.raster_settings(synthetic.orbit_camera(64, 48, 0.2).to(dev), 0, torch.zeros(3, device=dev), api.GaussianRasterizationSettings)
sc = synthetic.make_scene(900, "band", "precomp", -3.0, seed=2).to(dev)
print("filter", api.GaussianRasterizer(rs).visible_filter_indices(sc.means3D, sc.scales, sc.rotations)[1].numel())
torch.cuda.synchronize()
print("done")
