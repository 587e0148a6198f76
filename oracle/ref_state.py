"""Views into the reference rasterizer's private byte buffers (TEST INFRASTRUCTURE ONLY).

The reference carves typed arrays out of three byte tensors with `obtain(chunk, ptr, count, 128)`
(reference cuda_rasterizer/rasterizer_impl.h:21-27, rasterizer_impl.cu:155-194).  Alignment is
computed on absolute addresses; torch's caching allocator returns 512-byte aligned blocks, so the
offsets below (relative to the tensor start) are the same as the reference's.

Only arrays that precede the CUB temp storage are exposed (its size is CUB-version dependent):
  geom    : depths f32[P], clamped bool[3P], internal_radii i32[P], means2D f32[P,2],
            cov3D f32[P,6], conic_opacity f32[P,4], rgb f32[P,3], tiles_touched u32[P]
  binning : point_list u32[R], point_list_unsorted u32[R], keys u64[R], keys_unsorted u64[R]
  image   : accum_alpha f32[N], n_contrib u32[N], ranges u32[N,2]   (N = W*H; first #tiles rows used)
"""
from __future__ import annotations

import torch


def _carve(buf: torch.Tensor, specs):
    out = {}
    off = 0
    for name, dtype, shape in specs:
        off = (off + 127) & ~127
        n = 1
        for s in shape:
            n *= s
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        out[name] = buf[off:off + nbytes].view(dtype).view(*shape) if n > 0 else torch.empty(shape, dtype=dtype, device=buf.device)
        off += nbytes
    return out


def geom_views(geom: torch.Tensor, P: int):
    return _carve(geom, [
        ("depths", torch.float32, (P,)),
        ("clamped", torch.bool, (P, 3)),
        ("internal_radii", torch.int32, (P,)),
        ("means2D", torch.float32, (P, 2)),
        ("cov3D", torch.float32, (P, 6)),
        ("conic_opacity", torch.float32, (P, 4)),
        ("rgb", torch.float32, (P, 3)),
        ("tiles_touched", torch.int32, (P,)),
    ])


def binning_views(binning: torch.Tensor, R: int):
    return _carve(binning, [
        ("point_list", torch.int32, (R,)),
        ("point_list_unsorted", torch.int32, (R,)),
        ("keys", torch.int64, (R,)),
        ("keys_unsorted", torch.int64, (R,)),
    ])


def image_views(image: torch.Tensor, W: int, H: int):
    N = W * H
    return _carve(image, [
        ("accum_alpha", torch.float32, (N,)),
        ("n_contrib", torch.int32, (N,)),
        ("ranges", torch.int32, (N, 2)),
    ])
