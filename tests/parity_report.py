"""Standalone GPU report: stage-level parity against the reference's own CUDA build + timings.

Run on a B200 box:  python tests/parity_report.py [--quick] [--out gpurun_out/parity_report.json]
Not a pytest file (the asserting versions of these checks live in tests/test_gpu_*.py); this one
prints everything it sees so a single GPU call is enough to localise a mismatch.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
import traceback

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import parity_lib as pl  # noqa: E402
from workload import synthetic  # noqa: E402


def sort_check(report):
    from bloomscene_b200 import _C

    res = {}
    g = torch.Generator(device="cpu").manual_seed(3)
    for n, lo, hi, tag in [(1, 0, 32, "n1"), (31, 0, 32, "n31"), (4096, 0, 32, "n4096"), (4097, 0, 13, "n4097_13b"),
                           (100_003, 0, 32, "n100k"), (2_000_000, 0, 13, "n2M_13b"), (1_000_000, 0, 32, "n1M_32b"),
                           (300_000, 0, 3, "n300k_3b"), (50_000, 0, 17, "n50k_17b")]:
        keys = torch.randint(0, 2 ** 31 - 1, (n,), generator=g, dtype=torch.int64)
        if "n1M" in tag:  # depth-like: few distinct top bytes, many exact ties
            keys = (torch.rand(n, generator=g) * 3 + 0.2).view(torch.int32).long()
            keys[::7] = keys[0]
        keys = keys & ((1 << hi) - 1) if hi < 32 else keys
        k32 = keys.to(torch.int32).cuda()
        ko, vo = _C.sort_pairs(k32, None, lo, hi)
        ref_k, ref_i = torch.sort(keys.cuda(), stable=True)
        ok = bool(torch.equal(ko.long() & 0xFFFFFFFF, ref_k & 0xFFFFFFFF) and torch.equal(vo.long(), ref_i))
        res[tag] = ok
    report["sort"] = res
    print("sort:", res, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default="gpurun_out/parity_report.json")
    ap.add_argument("--no-timing", action="store_true")
    a = ap.parse_args()
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    dev = torch.device("cuda:0")
    report = {"gpu": torch.cuda.get_device_name(0), "ref_available": pl.reference() is not None}
    print(report, flush=True)

    try:
        sort_check(report)
    except Exception:
        report["sort_error"] = traceback.format_exc()
        print(report["sort_error"], flush=True)

    cases = [
        ("tiny17_sh0", dict(P=17, kind="object", color="sh0", mu=-3.0), (64, 48)),
        ("p1000_sh3", dict(P=1000, kind="object", color="sh3", mu=-3.5), (130, 70)),
        ("p20k_precomp", dict(P=20_000, kind="object", color="precomp", mu=-4.0), (256, 256)),
        ("p20k_sh1m16", dict(P=20_000, kind="object", color="sh1m16", mu=-4.0), (200, 120)),
        ("A_100k_sh0", dict(P=100_000, kind="object", color="sh0", mu=-4.0), (512, 512)),
    ]
    if not a.quick:
        cases.append(("C_1M_sh3", dict(P=1_000_000, kind="object", color="sh3", mu=-5.3), (1920, 1080)))
    report["cases"] = {}
    for name, sc, (W, H) in cases:
        try:
            scene = synthetic.make_scene(sc["P"], sc["kind"], sc["color"], sc["mu"], seed=0).to(dev)
            cam = synthetic.orbit_camera(W, H, 0.3).to(dev)
            bg = torch.tensor([0.2, 0.5, 0.7], device=dev)
            Wc, Wd = (t.to(dev) for t in synthetic.loss_weights(W, H))
            rep = {}
            if pl.reference() is not None:
                rep["stages"] = pl.compare_stages(scene, cam, bg)
                print(name, "stages", rep["stages"], flush=True)
                rep["autograd"] = pl.compare_autograd(scene, cam, bg, Wc, Wd)
                print(name, "autograd", rep["autograd"], flush=True)
            else:
                out = pl.run_autograd(pl.ours(), scene, cam, bg, Wc, Wd)
                rep["ours_only"] = {k: float(v.float().abs().mean().item()) for k, v in out["grads"].items() if v is not None}
                print(name, rep["ours_only"], flush=True)
            if not a.no_timing and sc["P"] >= 100_000:
                rep["time_ours"] = pl.time_fwd_bwd(pl.ours(), scene, cam, bg, Wc)
                print(name, "time ours", rep["time_ours"], flush=True)
                if pl.reference() is not None:
                    rep["time_ref"] = pl.time_fwd_bwd(pl.reference(), scene, cam, bg, Wc)
                    print(name, "time ref ", rep["time_ref"], flush=True)
            report["cases"][name] = rep
        except Exception:
            report["cases"][name] = {"error": traceback.format_exc()}
            print(name, "ERROR", report["cases"][name]["error"], flush=True)
        with open(a.out, "w") as f:
            json.dump(report, f, indent=1)
    print("done", flush=True)


if __name__ == "__main__":
    main()
