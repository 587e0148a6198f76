"""GPU parity of the fused steps either side of the rasterizer against the golden vectors of the
reference's own utils/loss.py and against the torch-op chains they replace (tests/ref_torch_ops.py).
Floating point: loss within 2e-6 absolute, gradients within 1e-4 relative L2 (the window sums are evaluated in a
different order than cuDNN's / ATen's convolutions)."""
import glob
import os

import numpy as np
import pytest
import torch

import parity_lib as pl
import ref_torch_ops as ref
from golden.make_golden_loss_inputs import make_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "loss_*.npz")))


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_l1_ssim_loss_vs_reference_golden(path):
    from bloomscene_b200.fused import l1_ssim_loss

    g = np.load(path)
    image, gt = make_inputs(int(g["C"]), int(g["H"]), int(g["W"]), int(g["seed"]))
    image = image.to(DEV).requires_grad_(True)
    loss = l1_ssim_loss(image, gt.to(DEV), float(g["lambda_dssim"]))
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) <= 2e-6
    assert pl.rel_l2(image.grad.cpu(), torch.from_numpy(g["grad"])) <= 1e-4


@pytest.mark.parametrize("shape", [(3, 512, 512), (3, 1080, 1920), (3, 17, 5), (1, 1, 1), (3, 33, 250)])
def test_l1_ssim_loss_vs_torch_ops(shape):
    from bloomscene_b200.fused import l1_ssim_loss

    C, H, W = shape
    image, gt = make_inputs(C, H, W, seed=H + W)
    a = image.to(DEV).requires_grad_(True)
    b = image.to(DEV).requires_grad_(True)
    gt = gt.to(DEV)
    got = l1_ssim_loss(a, gt, 0.2) * 3.0  # a non-unit upstream gradient
    want = ref.l1_ssim_reference(b, gt, 0.2) * 3.0
    got.backward()
    want.backward()
    assert abs(float(got) - float(want)) <= 6e-6
    assert pl.rel_l2(a.grad, b.grad) <= 1e-4


def _neural_inputs(N, K, seed, frac_masked=0.5):
    g = torch.Generator().manual_seed(seed)
    anchor, gs = torch.randn(N, 3, generator=g), torch.rand(N, 6, generator=g) + 0.05
    off = torch.randn(N, K, 3, generator=g)
    nop = torch.randn(N * K, 1, generator=g) + (0.0 if frac_masked == 0.5 else (10.0 if frac_masked == 0.0 else -10.0))
    nop[::7] = 0.0  # exactly zero is masked out (neural_opacity > 0)
    col, sr = torch.rand(N * K, 3, generator=g), torch.randn(N * K, 7, generator=g)
    return [t.to(DEV) for t in (anchor, gs, off, nop, col, sr)]


@pytest.mark.parametrize("case", [(5000, 10, 0.5), (1, 1, 0.5), (333, 7, 0.0), (129, 10, 1.0), (40_000, 10, 0.5), (64, 32, 0.5)])
def test_neural_gaussians_vs_torch_ops(case):
    from bloomscene_b200.fused import neural_gaussians

    N, K, fm = case
    ins = _neural_inputs(N, K, seed=N + K, frac_masked=fm)
    a = [t.clone().requires_grad_(True) for t in ins]
    b = [t.clone().requires_grad_(True) for t in ins]
    got = neural_gaussians(*a)
    want = ref.neural_gaussians_reference(*b)
    assert torch.equal(got[5], want[5])
    m = int(want[5].sum())
    names = ["xyz", "color", "opacity", "scaling", "rot"]
    for n, x, y in zip(names, got[:5], want[:5]):
        assert x.shape == y.shape == (m, y.shape[1]), n
        assert torch.allclose(x, y, rtol=2e-6, atol=2e-7), n
    if m == 0:
        return
    gen = torch.Generator().manual_seed(9)
    ws = [torch.randn(t.shape, generator=gen).to(DEV) for t in want[:5]]
    sum((x * w).sum() for x, w in zip(got[:5], ws)).backward()
    sum((y * w).sum() for y, w in zip(want[:5], ws)).backward()
    for n, x, y in zip(["anchor", "grid_scaling", "grid_offsets", "neural_opacity", "color", "scale_rot"], a, b):
        assert x.grad.shape == y.grad.shape, n
        assert pl.rel_l2(x.grad, y.grad) <= 1e-5, (n, pl.rel_l2(x.grad, y.grad))


def test_neural_gaussians_feed_the_rasterizer():
    """The fused epilogue writes the rasterizer's input layout: its outputs go straight into GaussianRasterizer."""
    from bloomscene_b200.fused import l1_ssim_loss, neural_gaussians
    from workload import synthetic

    api = pl.ours()
    N, K = 3000, 10
    anchor, gs, off, nop, col, sr = _neural_inputs(N, K, seed=3)
    anchor = (anchor * 0.5).requires_grad_(True)
    gs = (gs * 0.02).requires_grad_(True)
    xyz, color, opacity, scaling, rot, mask = neural_gaussians(anchor, gs, off * 0.5, torch.sigmoid(nop) * (nop > 0), col, sr)
    cam = synthetic.orbit_camera(128, 96, 0.3).to(DEV)
    st = synthetic.raster_settings(cam, 0, torch.zeros(3, device=DEV), api.GaussianRasterizationSettings)
    img, radii, depth = api.GaussianRasterizer(st)(means3D=xyz, means2D=torch.zeros_like(xyz, requires_grad=True), opacities=opacity,
                                                   colors_precomp=color, scales=scaling, rotations=rot)
    gt = torch.rand(3, 96, 128, device=DEV)
    loss = l1_ssim_loss(img, gt, 0.2) + 0.01 * scaling.prod(dim=1).mean()  # bloomscene.py:284-290
    loss.backward()
    assert torch.isfinite(anchor.grad).all() and anchor.grad.abs().sum() > 0 and gs.grad.abs().sum() > 0
    assert int((radii > 0).sum()) > 100
