"""Synthetic scenes and cameras for the parity tests and the benchmark.

Everything is generated on the CPU from a seeded `torch.Generator` so that the oracle, the reference
CUDA build and the B200-native library all see identical bits, then moved to the target device.

Camera maths follows BloomScene: `getWorld2View2` / `getProjectionMatrix` (reference
utils/graphics.py:43-77) and the row-vector storage of scene/cameras.py:59-62
(`viewmatrix = W2C^T`, `projmatrix = W2C^T @ Proj^T`, `campos = inverse(viewmatrix)[3, :3]`),
with BloomScene's intrinsics scaled with the width (focal = 582.69 px at 512 px, arguments.py:106).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np
import torch

FOCAL_512 = 5.8269e02
ZNEAR, ZFAR = 0.01, 100.0


@dataclass
class Camera:
    image_width: int
    image_height: int
    tanfovx: float
    tanfovy: float
    viewmatrix: torch.Tensor  # [4,4] = W2C^T
    projmatrix: torch.Tensor  # [4,4] = W2C^T @ Proj^T
    campos: torch.Tensor  # [3]

    def to(self, device) -> "Camera":
        return Camera(self.image_width, self.image_height, self.tanfovx, self.tanfovy,
                      self.viewmatrix.to(device), self.projmatrix.to(device), self.campos.to(device))


def world2view(R_c2w: np.ndarray, t_w2c: np.ndarray) -> np.ndarray:
    """reference utils/graphics.py:43-54 with translate=0, scale=1."""
    Rt = np.zeros((4, 4))
    Rt[:3, :3] = R_c2w.transpose()
    Rt[:3, 3] = t_w2c
    Rt[3, 3] = 1.0
    C2W = np.linalg.inv(Rt)
    Rt = np.linalg.inv(C2W)
    return np.float32(Rt)


def projection_matrix(znear: float, zfar: float, fovX: float, fovY: float) -> torch.Tensor:
    """reference utils/graphics.py:57-77."""
    tanHalfFovY = math.tan(fovY / 2)
    tanHalfFovX = math.tan(fovX / 2)
    top = tanHalfFovY * znear
    bottom = -top
    right = tanHalfFovX * znear
    left = -right
    P = torch.zeros(4, 4)
    z_sign = 1.0
    P[0, 0] = 2.0 * znear / (right - left)
    P[1, 1] = 2.0 * znear / (top - bottom)
    P[0, 2] = (right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[3, 2] = z_sign
    P[2, 2] = z_sign * zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def make_camera(W: int, H: int, R_c2w: np.ndarray, t_w2c: np.ndarray) -> Camera:
    focal = FOCAL_512 * W / 512.0
    tanfovx = W / (2.0 * focal)
    tanfovy = H / (2.0 * focal)
    fovx, fovy = 2 * math.atan(tanfovx), 2 * math.atan(tanfovy)
    view = torch.tensor(world2view(R_c2w, t_w2c)).transpose(0, 1).contiguous()
    proj = projection_matrix(ZNEAR, ZFAR, fovx, fovy).transpose(0, 1)
    full = (view.unsqueeze(0).bmm(proj.unsqueeze(0))).squeeze(0).contiguous()
    campos = view.inverse()[3, :3].contiguous()
    return Camera(W, H, math.tan(fovx * 0.5), math.tan(fovy * 0.5), view, full, campos)


def _rot_y(theta: float) -> np.ndarray:
    c, s = math.cos(theta), math.sin(theta)
    return np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])


def orbit_camera(W: int, H: int, yaw: float, distance: float = 3.2) -> Camera:
    """Camera on a circle of radius `distance` around the origin, looking at it ("object" scenes)."""
    R_c2w = _rot_y(yaw)
    # camera centre c = R_c2w @ (0,0,-d); W2C translation t = -R_c2w^T c = (0,0,d)
    return make_camera(W, H, R_c2w, np.array([0.0, 0.0, distance]))


def yaw_camera(W: int, H: int, yaw: float) -> Camera:
    """Camera at the origin yawing in place, like the rotate360 trajectory ("band" scenes)."""
    return make_camera(W, H, _rot_y(yaw), np.array([0.0, 0.0, 0.0]))


@dataclass
class Scene:
    means3D: torch.Tensor  # [P,3]
    scales: torch.Tensor  # [P,3]
    rotations: torch.Tensor  # [P,4] unit quaternions (the caller normalises, as BloomScene does)
    opacities: torch.Tensor  # [P,1]
    shs: Optional[torch.Tensor]  # [P,M,3]
    colors_precomp: Optional[torch.Tensor]  # [P,3]
    sh_degree: int

    @property
    def P(self) -> int:
        return self.means3D.shape[0]

    def to(self, device) -> "Scene":
        mv = lambda t: None if t is None else t.to(device)
        return Scene(mv(self.means3D), mv(self.scales), mv(self.rotations), mv(self.opacities), mv(self.shs),
                     mv(self.colors_precomp), self.sh_degree)

    def tensors(self) -> Dict[str, torch.Tensor]:
        d = {"means3D": self.means3D, "scales": self.scales, "rotations": self.rotations, "opacities": self.opacities}
        if self.shs is not None:
            d["shs"] = self.shs
        if self.colors_precomp is not None:
            d["colors_precomp"] = self.colors_precomp
        return d

    def param_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.tensors().values())


def make_scene(P: int, kind: str = "object", color: str = "sh3", log_scale_mean: float = -5.3, seed: int = 0,
               fovy_for_band: Optional[float] = None) -> Scene:
    """kind: "object" (means ~ U([-1,1]^3)) or "band" (ring around the origin).
    color: "sh0" | "sh1" | "sh2" | "sh3" (M = (deg+1)^2) | "sh3m16"-style "shD" with M=16 via "shDm16" | "precomp"."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    if kind == "object":
        means = torch.rand(P, 3, generator=g) * 2.0 - 1.0
    elif kind == "band":
        az = torch.rand(P, generator=g) * (2 * math.pi)
        half = 1.15 * (fovy_for_band if fovy_for_band is not None else 2 * math.atan(512 / (2 * FOCAL_512))) / 2
        el = (torch.rand(P, generator=g) * 2.0 - 1.0) * half
        r = 1.5 + torch.rand(P, generator=g) * 2.5
        means = torch.stack([r * torch.cos(el) * torch.sin(az), r * torch.sin(el), r * torch.cos(el) * torch.cos(az)], dim=1)
    else:
        raise ValueError(kind)
    scales = torch.exp(log_scale_mean + 0.5 * torch.randn(P, 3, generator=g))
    q = torch.randn(P, 4, generator=g)
    rotations = q / q.norm(dim=1, keepdim=True)
    opacities = torch.sigmoid(1.5 * torch.randn(P, 1, generator=g))
    shs = colors = None
    deg = 0
    if color == "precomp":
        colors = torch.rand(P, 3, generator=g)
    else:
        spec = color.lower()
        deg = int(spec[2])
        M = 16 if spec.endswith("m16") else (deg + 1) ** 2
        shs = torch.randn(P, M, 3, generator=g) * 0.1
        shs[:, 0, :] = torch.randn(P, 3, generator=g) / 0.564
    return Scene(means.contiguous(), scales.contiguous(), rotations.contiguous(), opacities.contiguous(),
                 None if shs is None else shs.contiguous(), colors, deg)


# ---- the configurations of BASELINE.json ---------------------------------------------------------

CONFIGS = {
    # name: (P, kind, color, W, H, log-scale mean, n_views)
    "A": dict(P=100_000, kind="object", color="sh0", W=512, H=512, mu=-4.0, views=1),
    "B": dict(P=500_000, kind="band", color="precomp", W=512, H=512, mu=-4.3, views=120),
    "C": dict(P=1_000_000, kind="object", color="sh3", W=1920, H=1080, mu=-5.3, views=1),
    "D": dict(P=3_000_000, kind="band", color="sh3", W=3840, H=2160, mu=-6.2, views=120),
    "E": dict(P=1_000_000, kind="object", color="sh3", W=1920, H=1080, mu=-5.3, views=64),
}


def config_scene(name: str, seed: int = 0, P: Optional[int] = None) -> Scene:
    c = CONFIGS[name]
    fovy = 2 * math.atan(c["H"] / (2 * FOCAL_512 * c["W"] / 512.0))
    return make_scene(P if P is not None else c["P"], c["kind"], c["color"], c["mu"], seed, fovy_for_band=fovy)


def config_cameras(name: str, n_views: Optional[int] = None) -> List[Camera]:
    c = CONFIGS[name]
    n = n_views if n_views is not None else c["views"]
    total = c["views"]
    if c["kind"] == "object":
        return [orbit_camera(c["W"], c["H"], 2 * math.pi * k / max(total, 1)) for k in range(n)]
    return [yaw_camera(c["W"], c["H"], math.radians(3.0) * k) for k in range(n)]


def loss_weights(W: int, H: int, seed: int = 1):
    """Fixed dense weights for the backward seed: loss = (color*Wc).sum() + (depth*Wd).sum()."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.rand(3, H, W, generator=g), torch.rand(1, H, W, generator=g)


def raster_settings(cam: Camera, sh_degree: int, bg: torch.Tensor, settings_cls, scale_modifier: float = 1.0,
                    prefiltered: bool = False, debug: bool = False):
    return settings_cls(image_height=cam.image_height, image_width=cam.image_width, tanfovx=cam.tanfovx,
                        tanfovy=cam.tanfovy, bg=bg, scale_modifier=scale_modifier, viewmatrix=cam.viewmatrix,
                        projmatrix=cam.projmatrix, sh_degree=sh_degree, campos=cam.campos, prefiltered=prefiltered,
                        debug=debug)
