"""bloomscene_b200 — B200-native (sm_100a) differentiable Gaussian rasterizer for BloomScene.

The only product here is the hot path of BloomScene's `submodules/depth-diff-gaussian-rasterization`:
hand-written CUDA kernels behind a C-ABI (include/bloomrast.h, libbloomrast.so), a thin torch
extension (`_C`) and a Python surface identical to the reference package.  There is NO CPU or
PyTorch fallback: importing this package fails loudly if the native extension is not built.
"""
from __future__ import annotations

import importlib
import os

__version__ = "0.1.0"

_HERE = os.path.dirname(os.path.abspath(__file__))


def _load_native():
    import torch  # noqa: F401  (libtorch symbols must be loaded before the extension)

    try:
        return importlib.import_module("bloomscene_b200._C")
    except ImportError as e:  # pragma: no cover - exercised only on a broken install
        raise ImportError(
            "bloomscene_b200: the native extension bloomscene_b200/_C.so (and libbloomrast.so) is not built. "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` or `python bloomscene_b200/build.py`. "
            "There is no CPU / PyTorch fallback for the rasterizer. Original error: %s" % (e,)
        ) from e


_C = _load_native()

from .rasterizer import GaussianRasterizationSettings, bind, cpu_deep_copy_tuple  # noqa: E402

_api = bind(_C)
GaussianRasterizer = _api.GaussianRasterizer
rasterize_gaussians = _api.rasterize_gaussians
_RasterizeGaussians = _api._RasterizeGaussians
render_views = _api.render_views

__all__ = [
    "GaussianRasterizationSettings",
    "GaussianRasterizer",
    "rasterize_gaussians",
    "render_views",
    "_RasterizeGaussians",
    "cpu_deep_copy_tuple",
    "bind",
    "_C",
]
