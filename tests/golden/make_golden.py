"""Generate tests/golden/*.npz from the reference's OWN CUDA rasterizer (oracle/_ref/_ref_C.so).

Run on a B200 box (the reference extension needs a GPU):
    python tests/golden/make_golden.py --out gpurun_out/golden
then copy the files into tests/golden/.  Each file holds the inputs (so the fixtures do not depend
on RNG reproducibility across torch versions), the public outputs, the reference's internal
integer state (keys, sorted point list, tile ranges, n_contrib) and all gradients for the loss
    loss = (color * Wc).sum() + (depth * Wd).sum().
These vectors pin the CPU oracle (tests/test_oracle_golden.py, runs without a GPU) and are a second,
box-independent check of the CUDA library (tests/test_gpu_golden.py).
"""
from __future__ import annotations

import argparse
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import parity_lib as pl  # noqa: E402
from workload import synthetic  # noqa: E402
from oracle import ref_state  # noqa: E402


def cov3d_from_scale_rot(scales, rots, mod=1.0):
    """Sigma = (S R)^T (S R) packed as the 6 upper-triangular entries (reference forward.cu:118-152)."""
    s = scales * mod
    r, x, y, z = rots[:, 0], rots[:, 1], rots[:, 2], rots[:, 3]
    # glm column-major R as built in the reference; M = S * R, Sigma = M^T M
    R = torch.stack([
        torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)], -1),
        torch.stack([2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)], -1),
        torch.stack([2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], -1)], 1)  # [P, col, row]
    Rm = R.transpose(1, 2)  # maths matrix [P, row, col]
    M = torch.diag_embed(s) @ Rm
    Sigma = M.transpose(1, 2) @ M
    return torch.stack([Sigma[:, 0, 0], Sigma[:, 0, 1], Sigma[:, 0, 2], Sigma[:, 1, 1], Sigma[:, 1, 2], Sigma[:, 2, 2]], 1).contiguous()


def cases():
    mk = synthetic.make_scene
    out = []
    out.append(("sh3_ragged", mk(2000, "object", "sh3", -3.6, seed=1), synthetic.orbit_camera(130, 70, 0.3), (0.0, 0.0, 0.0), 1.0, False))
    out.append(("precomp_bg_mod", mk(3000, "object", "precomp", -3.8, seed=2), synthetic.orbit_camera(200, 120, 1.1), (0.2, 0.5, 0.7), 1.3, False))
    out.append(("cov3d_sh0", mk(1500, "object", "sh0", -3.5, seed=3), synthetic.orbit_camera(96, 96, 2.0), (1.0, 1.0, 1.0), 1.0, True))
    out.append(("band_culled", mk(4000, "band", "precomp", -3.0, seed=4), synthetic.yaw_camera(64, 64, 0.5), (0.0, 0.0, 0.0), 1.0, False))
    out.append(("sh1_m16", mk(1000, "object", "sh1m16", -3.3, seed=5), synthetic.orbit_camera(80, 48, 4.0), (0.1, 0.1, 0.1), 1.0, False))
    # exact depth ties: every Gaussian duplicated -> exercises sort stability
    s = mk(300, "object", "sh2", -3.0, seed=6)
    dup = synthetic.Scene(*[None if t is None else torch.cat([t, t]).contiguous() for t in
                            (s.means3D, s.scales, s.rotations, s.opacities, s.shs, s.colors_precomp)], s.sh_degree)
    out.append(("depth_ties", dup, synthetic.orbit_camera(64, 64, 0.0), (0.0, 0.0, 0.0), 1.0, False))
    # one huge opaque Gaussian + small ones: whole-screen radius, 0.99 clamp, early termination
    s = mk(500, "object", "sh0", -3.0, seed=7)
    s.scales[0] = torch.tensor([0.8, 0.8, 0.8])
    s.means3D[0] = torch.tensor([0.0, 0.0, -1.5])
    s.opacities[:50] = 1.0
    out.append(("saturating", s, synthetic.orbit_camera(72, 40, 0.0), (0.3, 0.3, 0.3), 1.0, False))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/golden")
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    ref = pl.reference()
    assert ref is not None, "oracle/_ref/_ref_C.so missing: run python oracle/build_ref.py where /root/reference exists"
    dev = torch.device("cuda:0")
    for name, scene_cpu, cam_cpu, bg, mod, use_cov in cases():
        scene, cam = scene_cpu.to(dev), cam_cpu.to(dev)
        bgt = torch.tensor(bg, device=dev)
        W, H = cam.image_width, cam.image_height
        Wc, Wd = (t.to(dev) for t in synthetic.loss_weights(W, H, seed=11))
        cov = cov3d_from_scale_rot(scene_cpu.scales, scene_cpu.rotations).to(dev) if use_cov else None
        res = pl.run_autograd(ref, scene, cam, bgt, Wc, Wd, scale_modifier=mod, cov3D=cov)
        args = pl.forward_args(scene, cam, bgt, mod, cov)
        R, color, depth, radii, geom, binning, img = ref._C.rasterize_gaussians(*args)
        P = scene.P
        gv, iv = ref_state.geom_views(geom, P), ref_state.image_views(img, W, H)
        ntiles = ((W + 15) // 16) * ((H + 15) // 16)
        d = {
            "W": W, "H": H, "tanfovx": cam.tanfovx, "tanfovy": cam.tanfovy, "bg": np.array(bg, np.float32),
            "scale_modifier": np.float32(mod), "sh_degree": scene.sh_degree,
            "viewmatrix": cam_cpu.viewmatrix.numpy(), "projmatrix": cam_cpu.projmatrix.numpy(),
            "campos": cam_cpu.campos.numpy(),
            "means3D": scene_cpu.means3D.numpy(), "opacities": scene_cpu.opacities.numpy(),
            "Wc": Wc.cpu().numpy(), "Wd": Wd.cpu().numpy(),
            "num_rendered": R, "color": color.cpu().numpy(), "depth": depth.cpu().numpy(), "radii": radii.cpu().numpy(),
            "depths": gv["depths"].cpu().numpy(), "means2D": gv["means2D"].cpu().numpy(),
            "conic_opacity": gv["conic_opacity"].cpu().numpy(), "tiles_touched": gv["tiles_touched"].cpu().numpy(),
            "n_contrib": iv["n_contrib"].cpu().numpy(), "final_T": iv["accum_alpha"].cpu().numpy(),
            "ranges": iv["ranges"][:ntiles].cpu().numpy(),
        }
        if use_cov:
            d["cov3D_precomp"] = cov.cpu().numpy()
        else:
            d["scales"], d["rotations"] = scene_cpu.scales.numpy(), scene_cpu.rotations.numpy()
        if scene_cpu.shs is not None:
            d["shs"] = scene_cpu.shs.numpy()
            d["rgb"] = gv["rgb"].cpu().numpy()
        else:
            d["colors_precomp"] = scene_cpu.colors_precomp.numpy()
        if R > 0:
            bv = ref_state.binning_views(binning, R)
            d["point_list"] = bv["point_list"].cpu().numpy()
            d["keys"] = bv["keys"].cpu().numpy()
            d["keys_unsorted"] = bv["keys_unsorted"].cpu().numpy()
        for k, v in res["grads"].items():
            if v is not None:
                d["grad_" + k] = v.cpu().numpy()
        path = os.path.join(a.out, name + ".npz")
        np.savez_compressed(path, **d)
        print(name, "P", P, "R", R, "visible", int((radii > 0).sum()), os.path.getsize(path) // 1024, "KiB", flush=True)


if __name__ == "__main__":
    main()
