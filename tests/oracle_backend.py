"""A `_C`-shaped module backed by the CPU oracle — TESTS ONLY.

`bloomscene_b200.rasterizer.bind(OracleBackend())` yields the full Python API (autograd Function,
GaussianRasterizer) running on CPU tensors, so host-side logic (argument plumbing, gradient order,
view sharding over gloo) can be tested where there is no GPU.  The product never imports this."""
from __future__ import annotations

import numpy as np
import torch

from oracle import oracle as orc


def _np(t):
    if t is None or (hasattr(t, "numel") and t.numel() == 0):
        return None
    return t.detach().cpu().numpy()


class OracleBackend:
    def __init__(self, threads: int = 0):
        self.threads = threads
        self._live = {}
        self._next = 1

    def rasterize_gaussians(self, bg, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp,
                            viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degree, campos,
                            prefiltered, debug):
        if means3D.dim() != 2 or means3D.shape[1] != 3:
            raise RuntimeError("means3D must have dimensions (num_points, 3)")
        o = orc.Oracle(self.threads)
        R, color, depth, radii = o.forward(
            W=image_width, H=image_height, tanfovx=tan_fovx, tanfovy=tan_fovy, bg=_np(bg), viewmatrix=_np(viewmatrix),
            projmatrix=_np(projmatrix), campos=_np(campos), sh_degree=degree, means3D=_np(means3D),
            opacities=_np(opacity), shs=_np(sh), colors_precomp=_np(colors), scales=_np(scales),
            rotations=_np(rotations), cov3D_precomp=_np(cov3D_precomp), scale_modifier=scale_modifier)
        handle = self._next
        self._next += 1
        self._live[handle] = o
        geom = torch.tensor([handle], dtype=torch.int64)
        return (R, torch.from_numpy(color), torch.from_numpy(depth), torch.from_numpy(radii), geom,
                torch.empty(0, dtype=torch.uint8), torch.empty(0, dtype=torch.uint8))

    def rasterize_gaussians_backward(self, bg, means3D, radii, colors, scales, rotations, scale_modifier, cov3D_precomp,
                                     viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color, dL_dout_depth, sh,
                                     degree, campos, geomBuffer, R, binningBuffer, imageBuffer, debug):
        o = self._live.pop(int(geomBuffer[0].item()))
        g = o.backward(_np(dL_dout_color))
        t = lambda k: torch.from_numpy(g[k])
        return (t("dL_dmeans2D"), t("dL_dcolors"), t("dL_dopacity"), t("dL_dmeans3D"), t("dL_dcov3D"), t("dL_dsh"),
                t("dL_dscales"), t("dL_drotations"))

    def rasterize_gaussians_backward_depth(self, bg, means3D, radii, colors, scales, rotations, scale_modifier,
                                           cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color,
                                           dL_dout_depth, sh, degree, campos, geomBuffer, R, binningBuffer, imageBuffer,
                                           debug, out_depth):
        o = self._live.pop(int(geomBuffer[0].item()))
        g = o.backward(_np(dL_dout_color), _np(dL_dout_depth))
        t = lambda k: torch.from_numpy(g[k])
        return (t("dL_dmeans2D"), t("dL_dcolors"), t("dL_dopacity"), t("dL_dmeans3D"), t("dL_dcov3D"), t("dL_dsh"),
                t("dL_dscales"), t("dL_drotations"))

    def rasterize_aussians_filter(self, means3D, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix,
                                  projmatrix, tan_fovx, tan_fovy, image_height, image_width, prefiltered, debug):
        o = orc.Oracle(self.threads)
        s = _np(scales)
        radii = o.visible_filter(W=image_width, H=image_height, tanfovx=tan_fovx, tanfovy=tan_fovy,
                                 viewmatrix=_np(viewmatrix), projmatrix=_np(projmatrix), means3D=_np(means3D),
                                 scales=None if s is None else np.ascontiguousarray(s), rotations=_np(rotations),
                                 cov3D_precomp=_np(cov3D_precomp), scale_modifier=scale_modifier)
        return torch.from_numpy(radii)

    def mark_visible(self, means3D, viewmatrix, projmatrix):
        return torch.from_numpy(orc.Oracle().mark_visible(_np(means3D), _np(viewmatrix)))
