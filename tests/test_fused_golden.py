"""CPU: the torch restatements of the steps either side of the rasterizer (tests/ref_torch_ops.py) against golden
vectors produced by the reference's own utils/loss.py (tests/golden/loss_*.npz)."""
import glob
import os

import numpy as np
import pytest
import torch

import ref_torch_ops as ref
from golden.make_golden_loss_inputs import make_inputs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "loss_*.npz")))


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_l1_ssim_restatement_matches_reference_outputs(path):
    g = np.load(path)
    image, gt = make_inputs(int(g["C"]), int(g["H"]), int(g["W"]), int(g["seed"]))
    image.requires_grad_(True)
    loss = ref.l1_ssim_reference(image, gt, float(g["lambda_dssim"]))
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) <= 1e-7
    assert np.abs(image.grad.numpy() - g["grad"]).max() <= 1e-9 + 1e-6 * np.abs(g["grad"]).max()


def test_neural_gaussians_restatement_shapes_and_mask():
    g = torch.Generator().manual_seed(0)
    N, K = 50, 10
    anchor, gs = torch.randn(N, 3, generator=g), torch.rand(N, 6, generator=g)
    off, nop = torch.randn(N, K, 3, generator=g), torch.randn(N * K, 1, generator=g)
    col, sr = torch.rand(N * K, 3, generator=g), torch.randn(N * K, 7, generator=g)
    xyz, c, o, s, r, mask = ref.neural_gaussians_reference(anchor, gs, off, nop, col, sr)
    m = int(mask.sum())
    assert xyz.shape == (m, 3) and c.shape == (m, 3) and o.shape == (m, 1) and s.shape == (m, 3) and r.shape == (m, 4)
    assert torch.allclose(r.norm(dim=1), torch.ones(m), atol=1e-6) and (o > 0).all()
    first = int(torch.nonzero(mask)[0])
    n0, k0 = divmod(first, K)
    assert torch.allclose(xyz[0], anchor[n0] + off[n0, k0] * gs[n0, :3])
