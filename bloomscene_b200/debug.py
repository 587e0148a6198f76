"""Typed views into the library's private forward state (for stage-level parity tests and tools).

Offsets come from the C-ABI (`brs_state_layout`, include/bloomrast.h) so Python never duplicates
the layout arithmetic."""
from __future__ import annotations

import torch


def state_views(_C, geom: torch.Tensor, binning: torch.Tensor, image: torch.Tensor, P: int, R: int, W: int, H: int):
    lay = _C.state_layout(P, R, W, H)
    ntiles = ((W + 15) // 16) * ((H + 15) // 16)
    npix = W * H

    def view(buf, off, dtype, shape):
        n = 1
        for s in shape:
            n *= s
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        if n == 0:
            return torch.empty(shape, dtype=dtype, device=buf.device)
        return buf[off:off + nbytes].view(dtype).view(*shape)

    out = {
        "records": view(geom, lay["geom_records"], torch.float32, (P, 12)),
        "depth_key": view(geom, lay["geom_depth_key"], torch.int32, (P,)),
        "rect": view(geom, lay["geom_rect"], torch.int32, (P, 2)),
        "order": view(geom, lay["geom_order"], torch.int32, (P,)),
        "point_list": view(binning, lay["binning_point_list"], torch.int32, (R,)),
        "ranges": view(image, lay["image_ranges"], torch.int32, (ntiles, 2)),
        "final_T": view(image, lay["image_final_T"], torch.float32, (npix,)),
        "n_contrib": view(image, lay["image_n_contrib"], torch.int32, (npix,)),
    }
    rec = out["records"]
    out["means2D"] = rec[:, 0:2]
    out["cull_tau"] = rec[:, 2]
    out["conic_opacity"] = rec[:, 4:8]
    out["rgb"] = rec[:, 8:11]
    out["depths"] = rec[:, 11]
    return out
