"""Synthetic curve sets and cameras (seeded, CPU numpy) shared by tests and bench.

Cameras follow the reference's matrix conventions exactly
(utils/graphics_utils.py:37-70 getWorld2View2 / getProjectionMatrix,
scene/cameras.py:59-66): world_view_transform = W2C^T, full_proj_transform =
world_view_transform @ P^T, camera_center = inverse(world_view_transform)[3,:3],
znear 0.01, zfar 100.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch


@dataclass
class SynthCamera:
    image_width: int
    image_height: int
    FoVx: float
    FoVy: float
    world_view_transform: torch.Tensor   # (4,4) = W2C^T
    full_proj_transform: torch.Tensor    # (4,4)
    camera_center: torch.Tensor          # (3,)
    image_name: str = "synth"

    def to(self, device):
        return SynthCamera(self.image_width, self.image_height, self.FoVx, self.FoVy,
                           self.world_view_transform.to(device), self.full_proj_transform.to(device),
                           self.camera_center.to(device), self.image_name)


def projection_matrix(znear: float, zfar: float, fovx: float, fovy: float) -> torch.Tensor:
    ty = math.tan(fovy / 2)
    tx = math.tan(fovx / 2)
    top, right = ty * znear, tx * znear
    bottom, left = -top, -right
    P = torch.zeros(4, 4)
    P[0, 0] = 2.0 * znear / (right - left)
    P[1, 1] = 2.0 * znear / (top - bottom)
    P[0, 2] = (right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def look_at_camera(eye, target, roll, width, height, fovx_deg=50.0, name="synth") -> SynthCamera:
    eye = np.asarray(eye, dtype=np.float64)
    target = np.asarray(target, dtype=np.float64)
    fwd = target - eye
    fwd /= np.linalg.norm(fwd)
    up = np.array([0.0, 0.0, 1.0])
    if abs(fwd @ up) > 0.999:
        up = np.array([0.0, 1.0, 0.0])
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    cr, sr = math.cos(roll), math.sin(roll)
    right, down = cr * right + sr * down, -sr * right + cr * down
    # camera axes (x right, y down, z forward) as columns of the C2W rotation == reference's R
    Rc2w = np.stack([right, down, fwd], axis=1)
    T = -Rc2w.T @ eye                        # W2C translation
    Rt = np.zeros((4, 4))
    Rt[:3, :3] = Rc2w.T
    Rt[:3, 3] = T
    Rt[3, 3] = 1.0
    Rt = np.float32(Rt)
    fovx = math.radians(fovx_deg)
    fovy = 2.0 * math.atan(math.tan(fovx / 2) * height / width)
    wvt = torch.tensor(Rt).transpose(0, 1).contiguous()
    proj = projection_matrix(0.01, 100.0, fovx, fovy).transpose(0, 1)
    full = wvt.unsqueeze(0).bmm(proj.unsqueeze(0)).squeeze(0).contiguous()
    center = wvt.inverse()[3, :3].contiguous()
    return SynthCamera(width, height, fovx, fovy, wvt, full, center, name)


def random_cameras(n, width, height, seed=0, rmin=1.8, rmax=2.6, fovx_deg=50.0):
    rng = np.random.default_rng(seed + 1000)
    cams = []
    for i in range(n):
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        r = rng.uniform(rmin, rmax)
        roll = rng.uniform(-math.pi, math.pi)
        eye = np.array([0.5, 0.5, 0.5]) + r * d
        cams.append(look_at_camera(eye, [0.5, 0.5, 0.5], roll, width, height, fovx_deg, name=f"synth_{i:04d}"))
    return cams


def random_curves(B, seed=0, line_fraction=0.0):
    """Unit-cube curve set: (curve_points (B,4,3), width (B,1), opacity (B,1), is_bezier (B,))."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(0.05, 0.95, size=(B, 3))
    d = rng.normal(size=(B, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    ell = rng.uniform(0.05, 0.25, size=(B, 1))
    P0 = c - 0.5 * ell * d
    P3 = c + 0.5 * ell * d
    P1 = P0 + ell * d / 3 + rng.normal(size=(B, 3)) * (0.15 * ell)
    P2 = P0 + 2 * ell * d / 3 + rng.normal(size=(B, 3)) * (0.15 * ell)
    pts = torch.tensor(np.stack([P0, P1, P2, P3], axis=1), dtype=torch.float32)
    width = torch.full((B, 1), math.log(5e-3), dtype=torch.float32)
    opacity = torch.full((B, 1), math.log(0.6 / 0.4), dtype=torch.float32)
    is_bezier = torch.tensor(rng.uniform(size=B) >= line_fraction)
    return pts, width, opacity, is_bezier


def random_gaussians(P, seed=0, scale_lo=0.003, scale_hi=0.03):
    """Generic (non-curve) Gaussian cloud for rasterizer-only tests."""
    g = torch.Generator().manual_seed(seed)
    means = torch.rand(P, 3, generator=g) * 0.9 + 0.05
    scales = torch.rand(P, 3, generator=g) * (scale_hi - scale_lo) + scale_lo
    rots = torch.randn(P, 4, generator=g)
    rots = rots / rots.norm(dim=1, keepdim=True)
    opac = torch.rand(P, 1, generator=g) * 0.8 + 0.1
    colors = torch.rand(P, 1, generator=g)
    all_map = torch.randn(P, 4, generator=g)
    return means, scales, rots, opac, colors, all_map
