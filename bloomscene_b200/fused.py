"""The two steps either side of the rasterizer in BloomScene's training iteration, each as one fused CUDA kernel
per direction.  Opt-in: BloomScene keeps working unchanged without them.

  neural_gaussians(...)   <- reference gaussian_renderer/__init__.py:168-203 (mask / concat / boolean index /
                             split / sigmoid / normalise / anchor + offset * scale)
  l1_ssim_loss(...)       <- reference utils/loss.py:83-84, 91-135 and bloomscene.py:284-287

Both need the native extension (there is no PyTorch fallback here: the torch-op chains ARE the reference).
"""
from __future__ import annotations

import torch

from . import _C


class _L1Ssim(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, target, lambda_dssim):
        image_c, target_c = image.contiguous(), target.contiguous()
        dmaps, partial = _C.l1_ssim_forward(image_c, target_c)
        sums = partial.sum(dim=0)  # (sum of SSIM, sum of |x - y|), fixed summation order
        n = float(image.numel())
        ctx.save_for_backward(image_c, target_c, dmaps)
        ctx.lambda_dssim = float(lambda_dssim)
        return (1.0 - lambda_dssim) * (sums[1] / n) + lambda_dssim * (1.0 - sums[0] / n)

    @staticmethod
    def backward(ctx, grad_loss):
        image, target, dmaps = ctx.saved_tensors
        g = grad_loss.reshape(1).to(torch.float32).contiguous()
        return _C.l1_ssim_backward(image, target, dmaps, g, ctx.lambda_dssim), None, None


def l1_ssim_loss(image: torch.Tensor, target: torch.Tensor, lambda_dssim: float = 0.2) -> torch.Tensor:
    """(1 - lambda) * l1_loss(image, target) + lambda * (1 - ssim(image, target)) for [C,H,W] images — the photometric
    loss of reference bloomscene.py:284-287 (lambda_dssim = 0.2, arguments.py), one kernel forward, one backward.
    The gradient flows to `image` only (the target is data)."""
    if image.dim() != 3 or image.shape != target.shape:
        raise ValueError("l1_ssim_loss: image and target must both be [C,H,W]")
    return _L1Ssim.apply(image, target, lambda_dssim)


class _NeuralGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, grid_scaling, grid_offsets, neural_opacity, color, scale_rot):
        nk = neural_opacity.numel()
        offsets = grid_offsets.reshape(nk, 3)
        xyz, col, op, sc, rot, index, count = _C.neural_gaussians_forward(
            anchor, grid_scaling, offsets, neural_opacity.reshape(nk), color.reshape(nk, 3), scale_rot.reshape(nk, 7))
        m = int(count.item())  # the reference's boolean indexing synchronises here as well
        ctx.save_for_backward(grid_scaling, offsets, scale_rot.reshape(nk, 7), index)
        ctx.shapes = (grid_offsets.shape, neural_opacity.shape, color.shape, scale_rot.shape)
        mask = index >= 0
        ctx.mark_non_differentiable(mask)
        return xyz[:m], col[:m], op[:m], sc[:m], rot[:m], mask

    @staticmethod
    def backward(ctx, d_xyz, d_color, d_opacity, d_scaling, d_rot, _d_mask):
        grid_scaling, offsets, scale_rot, index = ctx.saved_tensors
        m = d_xyz.shape[0]
        z = lambda g, w: torch.zeros((m, w), dtype=torch.float32, device=index.device) if g is None else g
        d_anchor, d_gs, d_off, d_nop, d_col, d_sr = _C.neural_gaussians_backward(
            grid_scaling, offsets, scale_rot, index, z(d_xyz, 3), z(d_color, 3), z(d_opacity, 1), z(d_scaling, 3), z(d_rot, 4))
        so, sn, sc, ss = ctx.shapes
        return d_anchor, d_gs, d_off.reshape(so), d_nop.reshape(sn), d_col.reshape(sc), d_sr.reshape(ss)


def neural_gaussians(anchor, grid_scaling, grid_offsets, neural_opacity, color, scale_rot):
    """Epilogue of `generate_neural_gaussians` (reference gaussian_renderer/__init__.py:168-203) in one kernel.

    anchor [N,3], grid_scaling [N,6], grid_offsets [N,K,3] (or [N*K,3]), neural_opacity [N*K,1] (already multiplied by
    the binary grid mask), color [N*K,3], scale_rot [N*K,7].  Returns (xyz [M,3], color [M,3], opacity [M,1],
    scaling [M,3], rot [M,4], mask bool[N*K]) over the M rows with neural_opacity > 0, in row order — exactly what the
    reference hands to the rasterizer — with gradients to all six inputs."""
    return _NeuralGaussians.apply(anchor, grid_scaling, grid_offsets, neural_opacity, color, scale_rot)
