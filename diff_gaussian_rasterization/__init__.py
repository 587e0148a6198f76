"""Alias of `depth_diff_gaussian_rasterization` under the upstream Inria package name
(BASELINE.json's north_star calls the drop-in `diff_gaussian_rasterization`)."""
from depth_diff_gaussian_rasterization import *  # noqa: F401,F403
from depth_diff_gaussian_rasterization import (  # noqa: F401
    GaussianRasterizationSettings,
    GaussianRasterizer,
    _C,
    _RasterizeGaussians,
    rasterize_gaussians,
    render_views,
)
