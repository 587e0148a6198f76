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

from .rasterizer import GaussianRasterizationSettings, lane_pool, lane_streams
from workload.params import GaussianParams, shard_views
from workload.synthetic import Camera, raster_settings


def default_loss(color: torch.Tensor, depth: torch.Tensor, Wc: torch.Tensor, Wd: torch.Tensor) -> torch.Tensor:
    return (color * Wc).sum() + (depth * Wd).sum()


def view_sharded_step(params: GaussianParams, cameras: Sequence[Camera], bg: torch.Tensor, rasterizer_cls,
                      loss_fn: Callable[[torch.Tensor, torch.Tensor, int], torch.Tensor],
                      rank: int = 0, world: int = 1, group=None, allreduce: bool = True,
                      streams: int = 4, host_threads: bool = False, keep_means2D_grad: bool = False) -> Dict[str, object]:
    """Render this rank's slice of `cameras`, backpropagate `loss_fn(color, depth, view_index)`, sum the
    parameter gradients over ranks.  Returns the step loss (summed over all views), per-view radii
    counts and the number of views rendered locally.

    With the native rasterizer on a GPU the views are dealt round-robin onto `streams` CUDA streams: a
    view's preprocess / sort / binning kernels (small grids, latency- and bandwidth-bound) then run
    under another view's blend kernels (issue-bound), and the forward's one host wait for a view falls
    while the other stream still has a backward queued.  All streams add into the one gradient bucket:
    the kernel accumulates with float reductions (brs_grads.accumulate)."""
    params.zero_grad()
    mine = shard_views(len(cameras), rank, world)
    dev = params.flat.device
    visible = []
    # the native rasterizer adds parameter gradients straight into the bucket slices (grad_sink);
    # any other rasterizer (the reference build, the CPU stand-in of the tests) goes through autograd
    use_sink = getattr(rasterizer_cls, "supports_grad_sink", False)
    n_lanes = max(1, min(streams, len(mine))) if (use_sink and dev.type == "cuda") else 1
    sink = params.grads() if use_sink else None
    losses = [torch.zeros((), dtype=torch.float32, device=dev) for _ in range(n_lanes)]

    def one_view(vi: int, lane: int):
        cam = cameras[vi]
        settings = raster_settings(cam, params.sh_degree, bg, GaussianRasterizationSettings)
        rast = rasterizer_cls(settings, grad_sink=sink) if sink is not None else rasterizer_cls(raster_settings=settings)
        means2D = params.zero_means2D().detach().requires_grad_(True)  # fresh leaf over a shared zero buffer
        color, radii, depth = rast(means3D=params.tensors["means3D"], means2D=means2D,
                                   opacities=params.tensors["opacities"], shs=params.get("shs"),
                                   colors_precomp=params.get("colors_precomp"), scales=params.tensors["scales"],
                                   rotations=params.tensors["rotations"], cov3D_precomp=None)
        loss = loss_fn(color, depth, vi)
        loss.backward()
        losses[lane] += loss.detach()
        # per-view statistics for the caller's densification (reference scene/gaussian_model.py:742-759 reads
        # viewspace_points.grad[:, :2] and radii > 0 after every view)
        visible.append((vi, radii, means2D.grad if keep_means2D_grad else None))

    if n_lanes == 1:
        for vi in mine:
            one_view(vi, 0)
    else:
        cur = torch.cuda.current_stream(dev)
        lanes = lane_streams(dev, n_lanes)
        for st in lanes:
            st.wait_stream(cur)  # parameters, zeroed bucket
        if host_threads:
            # one host thread per stream: the native calls release the GIL, so one view's host wait and
            # kernel launches do not hold up the other streams' launches
            def drive(lane: int):
                with torch.cuda.stream(lanes[lane]):
                    for vi in mine[lane::n_lanes]:
                        one_view(vi, lane)

            pool = lane_pool(dev, n_lanes)
            for f in [pool.submit(drive, lane) for lane in range(n_lanes)]:
                f.result()
        else:
            for j, vi in enumerate(mine):
                with torch.cuda.stream(lanes[j % n_lanes]):
                    one_view(vi, j % n_lanes)
        for st in lanes:
            cur.wait_stream(st)
        for _, radii, m2 in visible:  # allocated on a lane stream, handed to the caller's stream
            radii.record_stream(cur)
            if m2 is not None:
                m2.record_stream(cur)
    loss_sum = losses[0]
    for extra in losses[1:]:
        loss_sum = loss_sum + extra
    if world > 1 and allreduce:
        import torch.distributed as dist

        dist.all_reduce(params.grad_bucket, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(loss_sum, op=dist.ReduceOp.SUM, group=group)
    return {"loss": loss_sum, "views": mine, "radii": [(vi, r) for vi, r, _ in visible],
            "means2D_grad": [(vi, m2) for vi, _, m2 in visible] if keep_means2D_grad else None}


def _slice_bounds(numel: int, rank: int, world: int):
    """Contiguous 1/world slice of a flat buffer owned by `rank` (last slice takes the remainder)."""
    per = (numel + world - 1) // world
    lo = min(numel, rank * per)
    return lo, min(numel, lo + per)


def upload_params(params: GaussianParams, host_flat: torch.Tensor, rank: int = 0, world: int = 1, group=None) -> int:
    """Host -> device refresh of the replicated parameter buffer for steps whose parameters live on the
    host (e.g. a host-side optimizer).  With N ranks every rank copies only ITS 1/N slice over its own
    PCIe link and the slices are exchanged over NVLink (all-gather), so the host traffic of the whole
    job is one copy of the parameters instead of N.  Returns the bytes this rank copied from the host."""
    n = params.flat.numel()
    with torch.no_grad():
        if world == 1:
            params.flat.copy_(host_flat, non_blocking=True)
            return n * 4
        import torch.distributed as dist

        lo, hi = _slice_bounds(n, rank, world)
        params.flat[lo:hi].copy_(host_flat[lo:hi], non_blocking=True)
        if n % world == 0:
            mine = params.flat[lo:hi]  # NCCL gathers in place (input = this rank's slot of the output)
            dist.all_gather_into_tensor(params.flat, mine if mine.is_cuda else mine.clone(), group=group)
        else:
            for r in range(world):
                a, b = _slice_bounds(n, r, world)
                if b > a:
                    dist.broadcast(params.flat[a:b], src=r, group=group)
        return (hi - lo) * 4


def download_grads(params: GaussianParams, host_grads: torch.Tensor, rank: int = 0, world: int = 1) -> int:
    """Device -> host copy of the (already all-reduced) gradient bucket: rank r writes slice r of the
    host buffer, so the N ranks of a node fill one host gradient with N parallel PCIe copies.
    Returns the bytes this rank copied."""
    lo, hi = _slice_bounds(params.grad_bucket.numel(), rank, world)
    host_grads[lo:hi].copy_(params.grad_bucket[lo:hi], non_blocking=True)
    return (hi - lo) * 4

