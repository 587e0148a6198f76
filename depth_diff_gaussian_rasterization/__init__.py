"""Drop-in for BloomScene's `depth_diff_gaussian_rasterization` (import name used at
gaussian_renderer/__init__.py:16 of the reference).  Same public names, backed by the B200-native
library in bloomscene_b200; put this repository root on PYTHONPATH instead of installing the
reference submodule."""
from bloomscene_b200 import (  # noqa: F401
    GaussianRasterizationSettings,
    GaussianRasterizer,
    _C,
    _RasterizeGaussians,
    cpu_deep_copy_tuple,
    rasterize_gaussians,
    render_views,
)
