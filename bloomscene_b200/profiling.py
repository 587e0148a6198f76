"""Per-stage timing and work counts of the native pipeline (measurement helper for bench.py).

Uses the library's own hooks (include/bloomrast.h): brs_stage_timing / brs_stage_times bracket every
stage with CUDA events on the launching stream, brs_count_pairs counts the (pixel, instance) pairs
E, C, E_b under the reference's per-pixel semantics (the algorithmic work unit of SURVEY.md §8d)."""
from __future__ import annotations

from typing import Dict, Sequence

import torch

from .debug import state_views


def profile_views(api, params, cameras: Sequence, bg: torch.Tensor, dL_dcolor: torch.Tensor, rank: int = 0,
                  world: int = 1, max_views: int = 8, warm: int = 1) -> Dict[str, object]:
    """Raw-binding forward + backward over up to `max_views` of this rank's views with stage timing on.
    Returns {"ms": per-stage milliseconds per view, "stats": per-view averages of V, R, E, C, Eb}."""
    _C = api._C
    e = torch.Tensor([])
    t = {k: v.detach() for k, v in params.tensors.items()}
    shs = t.get("shs", e)
    cols = t.get("colors_precomp", e)
    views = list(range(rank, len(cameras), world))[:max_views]
    agg = {"V": 0, "R": 0, "R1": 0, "E": 0, "C": 0, "Eb": 0}

    def one(cam, count: bool):
        R, color, depth, radii, geom, binning, img = _C.rasterize_gaussians(
            bg, t["means3D"], cols, t["opacities"], t["scales"], t["rotations"], 1.0, e, cam.viewmatrix,
            cam.projmatrix, cam.tanfovx, cam.tanfovy, cam.image_height, cam.image_width, shs, params.sh_degree,
            cam.campos, False, False)
        _C.rasterize_gaussians_backward(
            bg, t["means3D"], radii, cols, t["scales"], t["rotations"], 1.0, e, cam.viewmatrix, cam.projmatrix,
            cam.tanfovx, cam.tanfovy, dL_dcolor, e, shs, params.sh_degree, cam.campos, geom, R, binning, img, False)
        if count:
            E, C, Eb = _C.count_pairs(geom, binning, img, params.P, R, cam.image_width, cam.image_height)
            agg["V"] += int((radii > 0).sum().item())
            agg["R"] += R
            rect = state_views(_C, geom, binning, img, params.P, R, cam.image_width, cam.image_height)["rect"].long()
            x0, x1, y0, y1 = rect[:, 0] & 0xFFFF, rect[:, 0] >> 16, rect[:, 1] & 0xFFFF, rect[:, 1] >> 16
            live = (x1 > x0) & (y1 > y0)
            cells = ((((x1 - 1) >> 3) + 1) - (x0 >> 3)) * ((((y1 - 1) >> 3) + 1) - (y0 >> 3))
            agg["R1"] += int(cells[live].sum().item())
            agg["E"] += E
            agg["C"] += C
            agg["Eb"] += Eb

    for _ in range(warm):
        one(cameras[views[0]], False)
    torch.cuda.synchronize()
    _C.stage_times()  # clear
    _C.stage_timing(True)
    for vi in views:
        one(cameras[vi], False)
    times = _C.stage_times()
    _C.stage_timing(False)
    for vi in views:
        one(cameras[vi], True)
    n = len(views)
    ms = {k: v[0] / n for k, v in times.items()}
    stats = {k: v / n for k, v in agg.items()}
    stats["views_profiled"] = n
    return {"ms": ms, "stats": stats}
