"""View-sharded training step: the one place the hot path meets a collective.

Views are independent given the Gaussian set, so the path shards by camera view: every rank holds
the full (replicated) Gaussian parameters, renders views `rank, rank+N, ...` of the batch with the
rasterizer, and accumulates per-Gaussian parameter gradients locally.  The only exchange is ONE
sum-allreduce of the flat fp32 gradient bucket per step (NCCL over NVLink on GPUs, gloo in the CPU
tests).  `means2D` gradients and `radii` are per-view statistics (BloomScene's densification uses
per-view norms, reference scene/gaussian_model.py:756-759) and are not reduced.

The reference has no such mode (single process, one view per step: bloomscene.py:237-243); this is
the data-parallel scaling axis BASELINE.json names.  The rasterizer class is injected so host-side
logic can be tested on CPU with an oracle-backed stand-in (tests only).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import torch

from .rasterizer import GaussianRasterizationSettings
from .synthetic import Camera, Scene, raster_settings

_ORDER = ("means3D", "scales", "rotations", "opacities", "shs", "colors_precomp")


class GaussianParams:
    """Gaussian parameters packed in one flat fp32 buffer with a matching flat gradient bucket.

    Each parameter tensor is a leaf view into `flat`, and its `.grad` is preset to the matching view
    of `grad_bucket`, so autograd accumulates every view's gradients straight into the bucket and a
    single allreduce covers all parameters ((44 + 12 M) bytes per Gaussian)."""

    def __init__(self, scene: Scene):
        tensors = {k: v for k, v in scene.tensors().items()}
        self.names = [n for n in _ORDER if n in tensors]
        self.sh_degree = scene.sh_degree
        dev = scene.means3D.device
        sizes = [tensors[n].numel() for n in self.names]
        self.flat = torch.empty(sum(sizes), dtype=torch.float32, device=dev)
        self.grad_bucket = torch.zeros_like(self.flat)
        self.tensors: Dict[str, torch.Tensor] = {}
        off = 0
        for n, sz in zip(self.names, sizes):
            seg = self.flat[off:off + sz].view(tensors[n].shape)
            seg.copy_(tensors[n])
            seg.requires_grad_(True)
            seg.grad = self.grad_bucket[off:off + sz].view(tensors[n].shape)
            self.tensors[n] = seg
            off += sz

    @property
    def P(self) -> int:
        return self.tensors["means3D"].shape[0]

    def zero_grad(self):
        self.grad_bucket.zero_()

    def grads(self) -> Dict[str, torch.Tensor]:
        return {n: t.grad for n, t in self.tensors.items()}

    def get(self, name: str) -> Optional[torch.Tensor]:
        return self.tensors.get(name)


def shard_views(n_views: int, rank: int, world: int) -> List[int]:
    """Round-robin view assignment: rank r renders views r, r+N, r+2N, ..."""
    return list(range(rank, n_views, world))


def default_loss(color: torch.Tensor, depth: torch.Tensor, Wc: torch.Tensor, Wd: torch.Tensor) -> torch.Tensor:
    return (color * Wc).sum() + (depth * Wd).sum()


def view_sharded_step(params: GaussianParams, cameras: Sequence[Camera], bg: torch.Tensor, rasterizer_cls,
                      loss_fn: Callable[[torch.Tensor, torch.Tensor, int], torch.Tensor],
                      rank: int = 0, world: int = 1, group=None, allreduce: bool = True) -> Dict[str, object]:
    """Render this rank's slice of `cameras`, backpropagate `loss_fn(color, depth, view_index)`, sum the
    parameter gradients over ranks.  Returns the step loss (summed over all views), per-view radii
    counts and the number of views rendered locally."""
    params.zero_grad()
    mine = shard_views(len(cameras), rank, world)
    loss_sum = torch.zeros((), dtype=torch.float32, device=params.flat.device)
    visible = []
    # the native rasterizer adds parameter gradients straight into the bucket slices (grad_sink);
    # any other rasterizer (the reference build, the CPU stand-in of the tests) goes through autograd
    sink = params.grads() if getattr(rasterizer_cls, "supports_grad_sink", False) else None
    for vi in mine:
        cam = cameras[vi]
        settings = raster_settings(cam, params.sh_degree, bg, GaussianRasterizationSettings)
        rast = rasterizer_cls(settings, grad_sink=sink) if sink is not None else rasterizer_cls(raster_settings=settings)
        means2D = torch.zeros_like(params.tensors["means3D"], requires_grad=True)
        color, radii, depth = rast(means3D=params.tensors["means3D"], means2D=means2D,
                                   opacities=params.tensors["opacities"], shs=params.get("shs"),
                                   colors_precomp=params.get("colors_precomp"), scales=params.tensors["scales"],
                                   rotations=params.tensors["rotations"], cov3D_precomp=None)
        loss = loss_fn(color, depth, vi)
        loss.backward()
        loss_sum += loss.detach()
        visible.append((vi, radii))
    if world > 1 and allreduce:
        import torch.distributed as dist

        dist.all_reduce(params.grad_bucket, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(loss_sum, op=dist.ReduceOp.SUM, group=group)
    return {"loss": loss_sum, "views": mine, "radii": visible}
