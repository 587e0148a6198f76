#!/usr/bin/env python
"""Long seeded differential fuzz against the reference's own CUDA build (oracle/_ref), beyond what the test suite's
28-case sweep covers: P from 1 to 600 K (all three radix tile variants, V capacities from the high-water marks of the
PREVIOUS case of the same shape class), images from 1x1 to 700x500, every colour mode, band and object scenes,
forward modes AUTO (optimistic, with re-runs when the marks are stale) and EXACT, and every 5th case a stack of views
through brs_forward_views against per-view calls.

    python tools/fuzz_vs_reference.py [cases=300] [seed=1] > gpurun_out/fuzz.jsonl

A gradient that differs from the reference's by more than 1e-4 (relative L2) is adjudicated by the CPU oracle with
double-precision sums: it is a failure only if this library is the one farther from it.

One JSON line per failure, one per adjudicated case and a summary line; exit code 1 on any failure."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import parity_lib as pl  # noqa: E402
from workload import synthetic  # noqa: E402

DEV = "cuda:0"
INT_KEYS = ("radii_mismatch", "depth_bits_mismatch", "means2D_mismatch", "conic_opacity_mismatch", "tiles_touched_mismatch",
            "point_list_mismatch", "ranges_mismatch", "sorted_keys_mismatch", "n_contrib_mismatch", "final_T_mismatch")


NAMES = {"means3D": "dL_dmeans3D", "opacities": "dL_dopacity", "scales": "dL_dscales", "rotations": "dL_drotations",
         "shs": "dL_dsh", "colors_precomp": "dL_dcolors", "means2D": "dL_dmeans2D"}
adjudicated = []


def adjudicate(scene, cam, bg, Wc, Wd, mod):
    from oracle import oracle

    cpu = lambda t: None if t is None else t.cpu()
    sc = synthetic.Scene(cpu(scene.means3D), cpu(scene.scales), cpu(scene.rotations), cpu(scene.opacities), cpu(scene.shs),
                         cpu(scene.colors_precomp), scene.sh_degree)
    o = oracle.run_scene(sc, cam.to("cpu"), bg.cpu(), scale_modifier=mod)
    truth = o["oracle"].backward(Wc.cpu().numpy(), f64_sums=True)
    zero = torch.zeros_like(Wd)
    a = pl.run_autograd(pl.ours(), scene, cam, bg, Wc, zero, mod)
    b = pl.run_autograd(pl.reference(), scene, cam, bg, Wc, zero, mod)
    out = {}
    for k, g in a["grads"].items():
        if g is not None:
            t = torch.from_numpy(truth[NAMES[k]]).to(DEV).reshape(g.shape)
            out[k] = (float(pl.rel_l2(g, t)), float(pl.rel_l2(b["grads"][k], t)))
    return out


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    assert pl.reference() is not None, "oracle/_ref/_ref_C.so missing"
    api = pl.ours()
    rng = np.random.default_rng(seed)
    colors = ["sh0", "sh1", "sh2", "sh3", "sh1m16", "sh3", "precomp", "precomp"]
    shapes = [(1, 1), (7, 3), (16, 16), (17, 33), (130, 70), (256, 256), (320, 200), (512, 512), (700, 500)]
    failures, worst = 0, {"color": 0.0, "depth": 0.0, "grad": 0.0}
    for i in range(cases):
        big = i % 7 == 3
        P = int(rng.integers(100_000, 600_000)) if big else int(rng.integers(1, 30_000))
        W, H = shapes[int(rng.integers(0, len(shapes)))]
        kind = "band" if i % 3 == 2 else "object"
        color = colors[i % len(colors)]
        mu = float(rng.uniform(-5.5, -3.8)) if big else float(rng.uniform(-4.5, -1.5))
        mod = float(rng.uniform(0.4, 2.5))
        bg = torch.tensor(rng.uniform(0, 1, 3), dtype=torch.float32, device=DEV)
        scene = synthetic.make_scene(P, kind, color, mu, seed=int(rng.integers(0, 1 << 30))).to(DEV)
        yaw = float(rng.uniform(0, 6.28))
        cam = (synthetic.orbit_camera(W, H, yaw) if kind == "object" else synthetic.yaw_camera(W, H, yaw)).to(DEV)
        tag = {"case": i, "P": P, "W": W, "H": H, "kind": kind, "color": color, "mu": round(mu, 2), "mod": round(mod, 2)}
        try:
            rep = pl.compare_stages(scene, cam, bg, scale_modifier=mod)
            bad = {k: rep[k] for k in INT_KEYS if rep.get(k, 0) != 0}
            if rep["R_ours"] != rep["R_ref"]:
                bad["R"] = (rep["R_ours"], rep["R_ref"])
            if rep.get("color_maxabs", 0.0) > pl.COLOR_TOL or rep.get("depth_maxabs", 0.0) > pl.COLOR_TOL:
                bad["image"] = (rep.get("color_maxabs"), rep.get("depth_maxabs"))
            worst["color"] = max(worst["color"], rep.get("color_maxabs", 0.0))
            worst["depth"] = max(worst["depth"], rep.get("depth_maxabs", 0.0))
            if not big or i % 2 == 0:
                Wc, Wd = (t.to(DEV) for t in synthetic.loss_weights(W, H, seed=i))
                gr = pl.compare_autograd(scene, cam, bg, Wc, Wd, scale_modifier=mod)
                over = {k: v for k, v in gr.items() if k.startswith("grad_") and v > pl.GRAD_TOL}
                worst["grad"] = max([worst["grad"]] + [v for k, v in gr.items() if k.startswith("grad_")])
                if over:
                    # who is right?  The CPU oracle with its per-Gaussian sums accumulated in double is the referee
                    # (tests/test_gpu_parity.py::test_gradients_of_screen_filling_gaussians_...): the case only fails
                    # if this library is farther from it than the reference, or farther than the 1e-4 tolerance
                    verdict = adjudicate(scene, cam, bg, Wc, Wd, mod)
                    adjudicated.append({**tag, "over_tolerance_vs_reference": over, "vs_double_precision_sums": verdict})
                    for k, (e_ours, e_ref) in verdict.items():
                        if e_ours > pl.GRAD_TOL or e_ours > e_ref + 1e-6:
                            bad["grad_" + k] = {"ours_vs_f64": e_ours, "ref_vs_f64": e_ref}
            if i % 5 == 0:
                # a stack of views through brs_forward_views against per-view calls
                n = int(rng.integers(2, 7))
                cams = [(synthetic.orbit_camera(W, H, yaw + 0.4 * k) if kind == "object" else synthetic.yaw_camera(W, H, yaw + 0.4 * k)).to(DEV)
                        for k in range(n)]
                st = [synthetic.raster_settings(c, scene.sh_degree, bg, api.GaussianRasterizationSettings, scale_modifier=mod) for c in cams]
                kw = dict(shs=scene.shs, colors_precomp=scene.colors_precomp, scales=scene.scales, rotations=scene.rotations)
                a = api.render_views(st, scene.means3D, scene.opacities, keep_radii=True, stack=int(rng.integers(2, 5)),
                                     streams=int(rng.integers(1, 3)), **kw)
                b = api.render_views(st, scene.means3D, scene.opacities, keep_radii=True, streams=1, **kw)
                torch.cuda.synchronize()
                if not (torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and all(torch.equal(x, y) for x, y in zip(a[2], b[2]))):
                    bad["stack_vs_single"] = True
            if bad:
                failures += 1
                print(json.dumps({**tag, "failed": {k: (v if not isinstance(v, tuple) else list(v)) for k, v in bad.items()}}), flush=True)
        except Exception as e:  # noqa: BLE001
            failures += 1
            print(json.dumps({**tag, "exception": repr(e)[:400]}), flush=True)
        del scene
        if big:
            torch.cuda.empty_cache()
    for rec in adjudicated:
        print(json.dumps({"adjudicated": rec}), flush=True)
    print(json.dumps({"cases": cases, "seed": seed, "failures": failures, "gradient_cases_over_1e-4_vs_reference": len(adjudicated), "worst_color_maxabs": worst["color"],
                      "worst_depth_maxabs": worst["depth"], "worst_grad_rel_l2": worst["grad"],
                      "forward_stats": api._C.forward_stats(False)}), flush=True)
    sys.exit(1 if failures else 0)


if __name__ == "__main__":
    main()
