"""Python surface of the rasterizer — same names, arguments, return values and error behaviour as the
reference package `depth_diff_gaussian_rasterization` (reference:
submodules/depth-diff-gaussian-rasterization/depth_diff_gaussian_rasterization/__init__.py:17-250).

`bind(_C)` builds the public objects on top of a native module exposing the reference's four
binding functions (ext.cpp:15-20).  The product binds bloomscene_b200._C (the B200-native CUDA
library); tests bind the reference's own extension through the very same wrapper, so parity tests
exercise identical Python on both sides.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import NamedTuple

import torch
import torch.nn as nn


class GaussianRasterizationSettings(NamedTuple):
    # field order is part of the API (reference __init__.py:158-170)
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


def cpu_deep_copy_tuple(input_tuple):
    # reference __init__.py:17-19 (debug snapshots)
    copied_tensors = [item.cpu().clone() if isinstance(item, torch.Tensor) else item for item in input_tuple]
    return tuple(copied_tensors)


def bind(_C) -> SimpleNamespace:
    """Create (rasterize_gaussians, _RasterizeGaussians, GaussianRasterizer) bound to native module `_C`."""

    class _RasterizeGaussians(torch.autograd.Function):
        # reference __init__.py:44-156
        @staticmethod
        def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                    raster_settings):
            args = (
                raster_settings.bg,
                means3D,
                colors_precomp,
                opacities,
                scales,
                rotations,
                raster_settings.scale_modifier,
                cov3Ds_precomp,
                raster_settings.viewmatrix,
                raster_settings.projmatrix,
                raster_settings.tanfovx,
                raster_settings.tanfovy,
                raster_settings.image_height,
                raster_settings.image_width,
                sh,
                raster_settings.sh_degree,
                raster_settings.campos,
                raster_settings.prefiltered,
                raster_settings.debug,
            )
            if raster_settings.debug:
                cpu_args = cpu_deep_copy_tuple(args)  # copy them before they can be corrupted
                try:
                    num_rendered, color, depth, radii, geomBuffer, binningBuffer, imgBuffer = _C.rasterize_gaussians(*args)
                except Exception as ex:
                    torch.save(cpu_args, "snapshot_fw.dump")
                    print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                    raise ex
            else:
                num_rendered, color, depth, radii, geomBuffer, binningBuffer, imgBuffer = _C.rasterize_gaussians(*args)

            ctx.raster_settings = raster_settings
            ctx.num_rendered = num_rendered
            ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer,
                                  binningBuffer, imgBuffer)
            return color, radii, depth

        @staticmethod
        def backward(ctx, grad_out_color, grad_radii, grad_depth):
            num_rendered = ctx.num_rendered
            raster_settings = ctx.raster_settings
            (colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer, binningBuffer,
             imgBuffer) = ctx.saved_tensors

            args = (
                raster_settings.bg,
                means3D,
                radii,
                colors_precomp,
                scales,
                rotations,
                raster_settings.scale_modifier,
                cov3Ds_precomp,
                raster_settings.viewmatrix,
                raster_settings.projmatrix,
                raster_settings.tanfovx,
                raster_settings.tanfovy,
                grad_out_color,
                grad_depth,
                sh,
                raster_settings.sh_degree,
                raster_settings.campos,
                geomBuffer,
                num_rendered,
                binningBuffer,
                imgBuffer,
                raster_settings.debug,
            )
            if raster_settings.debug:
                cpu_args = cpu_deep_copy_tuple(args)
                try:
                    (grad_means2D, grad_colors_precomp, grad_opacities, grad_means3D, grad_cov3Ds_precomp, grad_sh,
                     grad_scales, grad_rotations) = _C.rasterize_gaussians_backward(*args)
                except Exception as ex:
                    torch.save(cpu_args, "snapshot_bw.dump")
                    print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                    raise ex
            else:
                (grad_means2D, grad_colors_precomp, grad_opacities, grad_means3D, grad_cov3Ds_precomp, grad_sh,
                 grad_scales, grad_rotations) = _C.rasterize_gaussians_backward(*args)

            # order of the autograd inputs (reference __init__.py:144-154); grad_radii / grad_depth carry nothing
            return (grad_means3D, grad_means2D, grad_sh, grad_colors_precomp, grad_opacities, grad_scales,
                    grad_rotations, grad_cov3Ds_precomp, None)

    def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                            raster_settings):
        # reference __init__.py:21-42
        return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                         cov3Ds_precomp, raster_settings)

    class GaussianRasterizer(nn.Module):
        # reference __init__.py:172-249
        def __init__(self, raster_settings):
            super().__init__()
            self.raster_settings = raster_settings

        def markVisible(self, positions):
            with torch.no_grad():
                raster_settings = self.raster_settings
                visible = _C.mark_visible(positions, raster_settings.viewmatrix, raster_settings.projmatrix)
            return visible

        def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                    cov3D_precomp=None):
            raster_settings = self.raster_settings

            if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
                raise Exception('Please provide excatly one of either SHs or precomputed colors!')

            if ((scales is None or rotations is None) and cov3D_precomp is None) or (
                    (scales is not None or rotations is not None) and cov3D_precomp is not None):
                raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')

            if shs is None:
                shs = torch.Tensor([])
            if colors_precomp is None:
                colors_precomp = torch.Tensor([])
            if scales is None:
                scales = torch.Tensor([])
            if rotations is None:
                rotations = torch.Tensor([])
            if cov3D_precomp is None:
                cov3D_precomp = torch.Tensor([])

            return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations,
                                       cov3D_precomp, raster_settings)

        def visible_filter(self, means3D, scales=None, rotations=None, cov3D_precomp=None):
            raster_settings = self.raster_settings

            if scales is None:
                scales = torch.Tensor([])
            if rotations is None:
                rotations = torch.Tensor([])
            if cov3D_precomp is None:
                cov3D_precomp = torch.Tensor([])

            with torch.no_grad():
                radii = _C.rasterize_aussians_filter(
                    means3D, scales, rotations, raster_settings.scale_modifier, cov3D_precomp,
                    raster_settings.viewmatrix, raster_settings.projmatrix, raster_settings.tanfovx,
                    raster_settings.tanfovy, raster_settings.image_height, raster_settings.image_width,
                    raster_settings.prefiltered, raster_settings.debug)
            return radii

    return SimpleNamespace(
        _C=_C,
        _RasterizeGaussians=_RasterizeGaussians,
        rasterize_gaussians=rasterize_gaussians,
        GaussianRasterizer=GaussianRasterizer,
        GaussianRasterizationSettings=GaussianRasterizationSettings,
    )
