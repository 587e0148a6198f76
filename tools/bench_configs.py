#!/usr/bin/env python
"""Per-config numbers of BASELINE.md §3 (configs A-D; config E is bench.py): ours next to the
reference's own CUDA rasterizer (oracle/_ref) on the same B200, same tensors, CUDA-event timed.

  A, C : single view, forward ms / backward ms through the raw bindings (median of 20 after 5 warm-ups)
         + bit-exact check of radii and colour max-abs against the reference on that view
  B, D : forward-only render of a slice of the rotate360 view list under no_grad, ms/view and views/s
         (wall time of the loop incl. the per-view host wait, device-synchronised at both ends)
         + radii / colour / depth parity against the reference on the first views

    python tools/bench_configs.py [--configs A,B,C,D] [--views 24] > gpurun_out/configs.jsonl

One JSON line per config.  Diagnostic / documentation tool; bench.py is the contract benchmark."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import parity_lib as pl  # noqa: E402
from workload import synthetic  # noqa: E402


def render_loop(api, scene, cams, bg, reps=1):
    """Forward-only rendering of `cams` the way BloomScene's render_video does (reference bloomscene.py:191-204)."""
    settings_cls = api.GaussianRasterizationSettings
    out = None
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.no_grad():
        for _ in range(reps):
            for cam in cams:
                rast = api.GaussianRasterizer(synthetic.raster_settings(cam, scene.sh_degree, bg, settings_cls))
                means2D = torch.zeros_like(scene.means3D)
                out = rast(means3D=scene.means3D, means2D=means2D, opacities=scene.opacities, shs=scene.shs,
                           colors_precomp=scene.colors_precomp, scales=scene.scales, rotations=scene.rotations)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return dt / (reps * len(cams)) * 1e3, out


def _median_pass(fn, reps, n_views):
    """ms per view of one pass over the view list: median over `reps` passes, each synchronised on both sides (a single
    pass can catch an allocator growth or a capacity re-run of a shape's first views)."""
    times, out = [], None
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        times.append((time.perf_counter() - t0) / n_views * 1e3)
    return sorted(times)[len(times) // 2], out


def render_batched(api, scene, cams, bg, streams=4, reps=5, host_threads=True):
    """The same frames through the forward-only batch entry point (render_views: views dealt onto CUDA streams)."""
    settings = [synthetic.raster_settings(cam, scene.sh_degree, bg, api.GaussianRasterizationSettings) for cam in cams]
    kw = dict(shs=scene.shs, colors_precomp=scene.colors_precomp, scales=scene.scales, rotations=scene.rotations, streams=streams,
              host_threads=host_threads)
    api.render_views(settings, scene.means3D, scene.opacities, **kw)  # warm-up: one full pass
    return _median_pass(lambda: api.render_views(settings, scene.means3D, scene.opacities, **kw), reps, len(cams))


def render_stacked(api, scene, cams, bg, stack, reps=5, streams=1):
    """The same frames `stack` views at a time as one pipeline on ONE stream (render_views(stack=...), brs_forward_views)."""
    settings = [synthetic.raster_settings(cam, scene.sh_degree, bg, api.GaussianRasterizationSettings) for cam in cams]
    kw = dict(shs=scene.shs, colors_precomp=scene.colors_precomp, scales=scene.scales, rotations=scene.rotations, stack=stack, streams=streams)
    api.render_views(settings, scene.means3D, scene.opacities, **kw)  # warm-up: one full pass (EXACT first, then the high-water marks settle)
    return _median_pass(lambda: api.render_views(settings, scene.means3D, scene.opacities, **kw), reps, len(cams))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="A,B,C,D")
    ap.add_argument("--views", type=int, default=24)
    ap.add_argument("--stacks", default="4,8", help="stack sizes of the one-pipeline path (configs B, D)")
    ap.add_argument("--no-ref", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    mine, ref = pl.ours(), pl.reference()
    for name in a.configs.split(","):
        cfg = synthetic.CONFIGS[name]
        scene = synthetic.config_scene(name).to(dev)
        bg = torch.zeros(3, device=dev)
        rec = {"config": name, "P": cfg["P"], "resolution": [cfg["W"], cfg["H"]], "color": cfg["color"], "gpu": torch.cuda.get_device_name(0)}
        if name in ("A", "C"):
            cam = synthetic.config_cameras(name, 1)[0].to(dev)
            Wc = synthetic.loss_weights(cfg["W"], cfg["H"])[0].to(dev)
            rec["ours"] = pl.time_fwd_bwd(mine, scene, cam, bg, Wc)
            if ref is not None:
                rec["reference_cuda"] = pl.time_fwd_bwd(ref, scene, cam, bg, Wc)
                o = mine._C.rasterize_gaussians(*pl.forward_args(scene, cam, bg))
                r = ref._C.rasterize_gaussians(*pl.forward_args(scene, cam, bg))
                rec["parity"] = {"R_equal": int(o[0]) == int(r[0]), "radii_mismatch": int((o[3] != r[3]).sum()),
                                 "color_maxabs": float((o[1] - r[1]).abs().max()), "depth_maxabs": float((o[2] - r[2]).abs().max())}
                rec["speedup_fwd_bwd"] = round((rec["reference_cuda"]["fwd_ms"] + rec["reference_cuda"]["bwd_ms"]) /
                                               (rec["ours"]["fwd_ms"] + rec["ours"]["bwd_ms"]), 2)
        else:
            cams = [c.to(dev) for c in synthetic.config_cameras(name, a.views)]
            render_loop(mine, scene, cams[:3], bg)  # warm-up
            ms, out = render_loop(mine, scene, cams, bg)
            rec["ours"] = {"fwd_ms_per_view": round(ms, 4), "views_per_s": round(1e3 / ms, 1), "views_timed": len(cams),
                           "visible_last_view": int((out[1] > 0).sum())}
            ms_b, outb = render_batched(mine, scene, cams, bg)
            rec["ours_render_views_4_streams"] = {"fwd_ms_per_view": round(ms_b, 4), "views_per_s": round(1e3 / ms_b, 1),
                                                  "last_frame_equal_to_loop": bool(torch.equal(outb[0][-1], out[0]))}
            ms_b1, _ = render_batched(mine, scene, cams, bg, host_threads=False)
            rec["ours_render_views_4_streams"]["fwd_ms_per_view_one_host_thread"] = round(ms_b1, 4)
            rec["ours_render_views_stacked_one_stream"] = {}
            for stack in [int(x) for x in a.stacks.split(",") if x]:
                ms_s, outs = render_stacked(mine, scene, cams, bg, stack)
                rec["ours_render_views_stacked_one_stream"][f"stack{stack}"] = {
                    "fwd_ms_per_view": round(ms_s, 4), "equal_to_multi_stream": bool(torch.equal(outs[0], outb[0]) and torch.equal(outs[1], outb[1]))}
                for lanes in (2, 3):
                    ms_l, outl = render_stacked(mine, scene, cams, bg, stack, streams=lanes)
                    rec["ours_render_views_stacked_one_stream"][f"stack{stack}"][f"fwd_ms_per_view_{lanes}_lanes"] = round(ms_l, 4)
                    rec["ours_render_views_stacked_one_stream"][f"stack{stack}"]["lanes_equal"] = bool(torch.equal(outl[0], outb[0]))
            if ref is not None and not a.no_ref:
                render_loop(ref, scene, cams[:3], bg)
                ms_r, _ = render_loop(ref, scene, cams, bg)
                rec["reference_cuda"] = {"fwd_ms_per_view": round(ms_r, 4), "views_per_s": round(1e3 / ms_r, 1)}
                rec["speedup_fwd"] = round(ms_r / ms, 2)
                rec["speedup_fwd_render_views"] = round(ms_r / ms_b, 2)
                worst = {"radii_mismatch": 0, "color_maxabs": 0.0, "depth_maxabs": 0.0, "R_equal": True}
                for cam in cams[:4]:
                    o = mine._C.rasterize_gaussians(*pl.forward_args(scene, cam, bg))
                    r = ref._C.rasterize_gaussians(*pl.forward_args(scene, cam, bg))
                    worst["R_equal"] &= int(o[0]) == int(r[0])
                    worst["radii_mismatch"] += int((o[3] != r[3]).sum())
                    worst["color_maxabs"] = max(worst["color_maxabs"], float((o[1] - r[1]).abs().max()))
                    worst["depth_maxabs"] = max(worst["depth_maxabs"], float((o[2] - r[2]).abs().max()))
                rec["parity_first_4_views"] = worst
        print(json.dumps(rec), flush=True)
        del scene
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
