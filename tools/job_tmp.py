import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import parity_lib as pl
from workload import synthetic
dev = torch.device("cuda:0")
api = pl.ours(); _C = api._C
scene = synthetic.config_scene("B").to(dev)
cams = [c.to(dev) for c in synthetic.config_cameras("B", 24)]
bg = torch.zeros(3, device=dev)
settings = [synthetic.raster_settings(cam, scene.sh_degree, bg, api.GaussianRasterizationSettings) for cam in cams]
kw = dict(colors_precomp=scene.colors_precomp, scales=scene.scales, rotations=scene.rotations)
for streams, ht in [(1, False), (2, False), (4, False), (4, True), (4, False)]:
    api.render_views(settings[:4], scene.means3D, scene.opacities, streams=streams, host_threads=ht, **kw)
    torch.cuda.synchronize(); _C.forward_stats(True)
    t0 = time.perf_counter()
    for _ in range(3):
        api.render_views(settings, scene.means3D, scene.opacities, streams=streams, host_threads=ht, **kw)
    torch.cuda.synchronize()
    print(streams, ht, round((time.perf_counter() - t0) / 72 * 1e3, 4), _C.forward_stats(False))
