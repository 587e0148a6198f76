"""Shared helpers for GPU parity tests: run an implementation (ours or the reference's own CUDA
extension from oracle/_ref) on a synthetic scene and compare every stage.

Parity levels (BASELINE.json north_star):
  bit-exact : radii, num_rendered, tile keys (tile id + depth bits), sorted point list, tile ranges,
              n_contrib, final_T
  <= 1e-5   : colour, depth (max abs)
  <= 1e-4   : every gradient tensor (relative L2)
"""
from __future__ import annotations

import os
import sys
from typing import Dict, Optional

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from workload import synthetic  # noqa: E402
from bloomscene_b200.rasterizer import GaussianRasterizationSettings, bind  # noqa: E402

COLOR_TOL = 1e-5
GRAD_TOL = 1e-4

_ref_api = None


def ours():
    import bloomscene_b200

    return bloomscene_b200._api


def reference():
    """The reference's own CUDA extension (oracle/_ref/_ref_C.so) under the same Python wrapper; None if absent."""
    global _ref_api
    if _ref_api is None:
        from oracle import build_ref

        mod = build_ref.load()
        if mod is None:
            return None
        _ref_api = bind(mod)
    return _ref_api


def forward_args(scene: synthetic.Scene, cam: synthetic.Camera, bg: torch.Tensor, scale_modifier=1.0, cov3D=None,
                 debug=False, prefiltered=False):
    e = torch.Tensor([])
    return (
        bg, scene.means3D,
        scene.colors_precomp if scene.colors_precomp is not None else e,
        scene.opacities,
        scene.scales if cov3D is None else e,
        scene.rotations if cov3D is None else e,
        scale_modifier,
        cov3D if cov3D is not None else e,
        cam.viewmatrix, cam.projmatrix, cam.tanfovx, cam.tanfovy, cam.image_height, cam.image_width,
        scene.shs if scene.shs is not None else e,
        scene.sh_degree, cam.campos, prefiltered, debug,
    )


def run_autograd(api, scene: synthetic.Scene, cam: synthetic.Camera, bg, Wc, Wd, scale_modifier=1.0, cov3D=None):
    """Forward + backward through the public Python API; returns outputs and gradients."""
    leaf = lambda t: None if t is None else t.detach().clone().requires_grad_(True)
    means3D, opac = leaf(scene.means3D), leaf(scene.opacities)
    scales, rots = (leaf(scene.scales), leaf(scene.rotations)) if cov3D is None else (None, None)
    cov = leaf(cov3D)
    shs, cols = leaf(scene.shs), leaf(scene.colors_precomp)
    means2D = torch.zeros_like(means3D, requires_grad=True)
    settings = synthetic.raster_settings(cam, scene.sh_degree, bg, GaussianRasterizationSettings,
                                         scale_modifier=scale_modifier)
    rast = api.GaussianRasterizer(raster_settings=settings)
    color, radii, depth = rast(means3D=means3D, means2D=means2D, opacities=opac, shs=shs, colors_precomp=cols,
                               scales=scales, rotations=rots, cov3D_precomp=cov)
    loss = (color * Wc).sum() + (depth * Wd).sum()
    loss.backward()
    out = {"color": color.detach(), "depth": depth.detach(), "radii": radii.detach()}
    grads = {"means3D": means3D.grad, "means2D": means2D.grad, "opacities": opac.grad}
    if scales is not None:
        grads["scales"], grads["rotations"] = scales.grad, rots.grad
    if cov is not None:
        grads["cov3D_precomp"] = cov.grad
    if shs is not None:
        grads["shs"] = shs.grad
    if cols is not None:
        grads["colors_precomp"] = cols.grad
    out["grads"] = grads
    return out


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double(), b.double()
    denom = b.norm().item()
    if denom == 0.0:
        return a.norm().item()
    return (a - b).norm().item() / denom


def compare_stages(scene, cam, bg, scale_modifier=1.0, cov3D=None) -> Dict[str, object]:
    """Raw-binding forward on both implementations and a stage-by-stage comparison (bit level)."""
    from bloomscene_b200.debug import state_views
    from oracle import ref_state

    mine, ref = ours(), reference()
    args = forward_args(scene, cam, bg, scale_modifier, cov3D)
    R0, col0, dep0, rad0, g0, b0, i0 = mine._C.rasterize_gaussians(*args)
    R1, col1, dep1, rad1, g1, b1, i1 = ref._C.rasterize_gaussians(*args)
    P, W, H = scene.P, cam.image_width, cam.image_height
    rep: Dict[str, object] = {"P": P, "R_ours": R0, "R_ref": R1}
    rep["radii_mismatch"] = int((rad0 != rad1).sum().item())
    rep["visible"] = int((rad1 > 0).sum().item())
    if P == 0:
        rep["color_maxabs"] = float((col0 - col1).abs().max().item()) if col0.numel() else 0.0
        return rep
    sv = state_views(mine._C, g0, b0, i0, P, R0, W, H)
    rg, ri = ref_state.geom_views(g1, P), ref_state.image_views(i1, W, H)
    vis = rad1 > 0
    # per-Gaussian intermediates (bitwise on visible Gaussians; -0.0 == 0.0 tolerated through float ==)
    rep["depth_bits_mismatch"] = int((sv["depth_key"][vis] != rg["depths"].view(torch.int32)[vis]).sum().item())
    rep["means2D_mismatch"] = int((sv["means2D"][vis] != rg["means2D"][vis]).any(dim=1).sum().item())
    rep["conic_opacity_mismatch"] = int((sv["conic_opacity"][vis] != rg["conic_opacity"][vis]).any(dim=1).sum().item())
    if scene.shs is not None:
        rep["rgb_maxabs"] = float((sv["rgb"][vis] - rg["rgb"][vis]).abs().max().item()) if vis.any() else 0.0
        rep["rgb_mismatch"] = int((sv["rgb"][vis] != rg["rgb"][vis]).any(dim=1).sum().item())
    rect = sv["rect"]
    tiles = ((rect[:, 0] >> 16) - (rect[:, 0] & 0xFFFF)) * ((rect[:, 1] >> 16) - (rect[:, 1] & 0xFFFF))
    rep["tiles_touched_mismatch"] = int((tiles != rg["tiles_touched"]).sum().item())
    if R0 == R1 and R1 > 0:
        rb = ref_state.binning_views(b1, R1)
        rep["point_list_mismatch"] = int((sv["point_list"] != rb["point_list"]).sum().item())
        # sorted 64-bit keys, reconstructed on our side from (tile of the range, depth bits of the id)
        ntiles = sv["ranges"].shape[0]
        rep["ranges_mismatch"] = int((sv["ranges"] != ri["ranges"][:ntiles]).any(dim=1).sum().item())
        starts = sv["ranges"][:, 0].long()
        ends = sv["ranges"][:, 1].long()
        tile_of = torch.repeat_interleave(torch.arange(ntiles, device=starts.device), (ends - starts).clamp(min=0))
        if tile_of.numel() == R1:
            dk = sv["depth_key"].long() & 0xFFFFFFFF
            keys = (tile_of << 32) | dk[sv["point_list"].long()]
            rep["sorted_keys_mismatch"] = int((keys != rb["keys"]).sum().item())
        else:
            rep["sorted_keys_mismatch"] = -1
    rep["n_contrib_mismatch"] = int((sv["n_contrib"] != ri["n_contrib"]).sum().item())
    rep["final_T_mismatch"] = int((sv["final_T"] != ri["accum_alpha"]).sum().item())
    rep["color_maxabs"] = float((col0 - col1).abs().max().item())
    rep["depth_maxabs"] = float((dep0 - dep1).abs().max().item())
    return rep


def compare_autograd(scene, cam, bg, Wc, Wd, scale_modifier=1.0, cov3D=None) -> Dict[str, float]:
    mine, ref = ours(), reference()
    a = run_autograd(mine, scene, cam, bg, Wc, Wd, scale_modifier, cov3D)
    b = run_autograd(ref, scene, cam, bg, Wc, Wd, scale_modifier, cov3D)
    b2 = run_autograd(ref, scene, cam, bg, Wc, Wd, scale_modifier, cov3D)  # reference run-to-run spread (float atomics)
    rep = {"color_maxabs": float((a["color"] - b["color"]).abs().max().item()) if a["color"].numel() else 0.0,
           "depth_maxabs": float((a["depth"] - b["depth"]).abs().max().item()) if a["depth"].numel() else 0.0,
           "radii_mismatch": int((a["radii"] != b["radii"]).sum().item())}
    for k in b["grads"]:
        ga, gb = a["grads"][k], b["grads"][k]
        if gb is None:
            rep["grad_" + k] = 0.0 if ga is None else float("nan")
            continue
        rep["grad_" + k] = rel_l2(ga, gb)
        rep["refspread_" + k] = rel_l2(b2["grads"][k], gb)
    return rep


def time_fwd_bwd(api, scene, cam, bg, Wc, iters=20, warmup=5):
    """Median CUDA-event times (ms) of the raw forward and backward bindings."""
    args = forward_args(scene, cam, bg)
    e = torch.Tensor([])
    fw, bw = [], []
    for it in range(warmup + iters):
        s0, s1, s2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        s0.record()
        R, color, depth, radii, geom, binning, img = api._C.rasterize_gaussians(*args)
        s1.record()
        api._C.rasterize_gaussians_backward(
            bg, scene.means3D, radii, scene.colors_precomp if scene.colors_precomp is not None else e,
            scene.scales, scene.rotations, 1.0, e, cam.viewmatrix, cam.projmatrix, cam.tanfovx, cam.tanfovy,
            Wc, e, scene.shs if scene.shs is not None else e, scene.sh_degree, cam.campos, geom, R, binning, img, False)
        s2.record()
        torch.cuda.synchronize()
        if it >= warmup:
            fw.append(s0.elapsed_time(s1))
            bw.append(s1.elapsed_time(s2))
    fw.sort()
    bw.sort()
    return {"fwd_ms": fw[len(fw) // 2], "bwd_ms": bw[len(bw) // 2], "R": int(R)}
