"""GPU: the CUDA library against golden vectors captured from the reference's own CUDA code
(tests/golden/*.npz, see tests/golden/make_golden.py).  Needs no reference build on the box."""
import os

import numpy as np
import pytest
import torch

import parity_lib as pl
from workload import synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
NAMES = ["sh3_ragged", "precomp_bg_mod", "cov3d_sh0", "band_culled", "sh1_m16", "depth_ties", "saturating"]
GRADS = ["means3D", "means2D", "opacities", "scales", "rotations", "shs", "colors_precomp", "cov3D_precomp"]


def _load(name):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    t = lambda k: torch.from_numpy(g[k]).to(DEV) if k in g else None
    scene = synthetic.Scene(t("means3D"), t("scales"), t("rotations"), t("opacities"), t("shs"), t("colors_precomp"),
                            int(g["sh_degree"]))
    cam = synthetic.Camera(int(g["W"]), int(g["H"]), float(g["tanfovx"]), float(g["tanfovy"]), t("viewmatrix"),
                           t("projmatrix"), t("campos"))
    return g, scene, cam, t("bg"), float(g["scale_modifier"]), t("cov3D_precomp")


@pytest.mark.parametrize("name", NAMES)
def test_against_golden(name):
    from bloomscene_b200.debug import state_views

    g, scene, cam, bg, mod, cov = _load(name)
    mine = pl.ours()
    if cov is not None:  # the scene object needs placeholders for the absent scale/rotation pair
        scene.scales = scene.rotations = None
    args = pl.forward_args(scene, cam, bg, mod, cov)
    R, color, depth, radii, geom, binning, img = mine._C.rasterize_gaussians(*args)
    W, H = cam.image_width, cam.image_height
    sv = state_views(mine._C, geom, binning, img, scene.P, R, W, H)
    eq = lambda a, k: np.array_equal(a.cpu().numpy().astype(np.int64), g[k].astype(np.int64))
    assert R == int(g["num_rendered"])
    assert eq(radii, "radii") and eq(sv["point_list"], "point_list") and eq(sv["ranges"], "ranges")
    assert eq(sv["n_contrib"], "n_contrib")
    vis = g["radii"] > 0
    assert np.array_equal(sv["depth_key"].cpu().numpy()[vis], g["depths"].view(np.int32)[vis])
    assert np.array_equal(sv["final_T"].cpu().numpy(), g["final_T"])
    assert np.array_equal(sv["means2D"].cpu().numpy()[vis], g["means2D"][vis])
    assert np.array_equal(sv["conic_opacity"].cpu().numpy()[vis], g["conic_opacity"][vis])
    # reconstructed 64-bit sorted keys == the reference's sorted keys
    ranges = sv["ranges"].long()
    tile_of = torch.repeat_interleave(torch.arange(ranges.shape[0], device=DEV), (ranges[:, 1] - ranges[:, 0]).clamp(min=0))
    keys = (tile_of << 32) | (sv["depth_key"].long() & 0xFFFFFFFF)[sv["point_list"].long()]
    assert np.array_equal(keys.cpu().numpy(), g["keys"])
    assert np.abs(color.cpu().numpy() - g["color"]).max() <= pl.COLOR_TOL
    assert np.abs(depth.cpu().numpy() - g["depth"]).max() <= pl.COLOR_TOL
    # gradients through the public autograd API
    Wc, Wd = torch.from_numpy(g["Wc"]).to(DEV), torch.from_numpy(g["Wd"]).to(DEV)
    if cov is not None:
        scene.scales = scene.rotations = torch.zeros(0, device=DEV)
    out = pl.run_autograd(mine, scene, cam, bg, Wc, Wd, scale_modifier=mod, cov3D=cov)
    for k in GRADS:
        if "grad_" + k in g:
            assert pl.rel_l2(out["grads"][k].cpu(), torch.from_numpy(g["grad_" + k])) <= pl.GRAD_TOL, k
