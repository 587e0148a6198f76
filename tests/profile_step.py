"""Tiny driver for ncu / timing runs: N forward+backward steps of one config on one implementation.

    python tests/profile_step.py --impl ours|ref --config C --iters 3
"""
from __future__ import annotations

import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import parity_lib as pl  # noqa: E402
from workload import synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", default="C")
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--view", type=int, default=0)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    api = pl.ours() if a.impl == "ours" else pl.reference()
    scene = synthetic.config_scene(a.config).to(dev)
    cam = synthetic.config_cameras(a.config, a.view + 1)[a.view].to(dev)
    bg = torch.zeros(3, device=dev)
    Wc, _ = synthetic.loss_weights(cam.image_width, cam.image_height)
    Wc = Wc.to(dev)
    t = pl.time_fwd_bwd(api, scene, cam, bg, Wc, iters=a.iters, warmup=1)
    print(a.impl, a.config, t)


if __name__ == "__main__":
    main()
