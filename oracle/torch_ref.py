"""TEST INFRASTRUCTURE ONLY — pure-PyTorch (CPU-capable, autograd) restatement of
the reference's Python-side math on the hot path. Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline leg may import this; the product never does.

Pinned against the reference itself: tests/golden/make_sampling_golden.py imports
/root/reference/utils/general_utils.py + scene/gaussian_curve_model.py (with stub
modules for the missing third-party imports) and stores its outputs under
tests/golden/; tests/test_oracle_sampling.py checks this file against them.

Restates:
  sample_curves          scene/gaussian_curve_model.py:58-60,70-89,180-198
  rot_to_quat            utils/general_utils.py:9-86 (rot_to_quat_batch)
  quaternion_to_matrix   pytorch3d.transforms (third-party, un-pinned; formula in SURVEY.md 8c)
  raster_inputs          scene/gaussian_curve_model.py:99-122 + gaussian_renderer/__init__.py:57-104
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def sample_t(n: int, device="cpu") -> torch.Tensor:
    return torch.linspace(0.5 / n, 1 - 0.5 / n, n, device=device)


def _curve_points_at(cp, is_bezier, t):
    """t: (n,1,1); cp: (B,4,3) -> (n,B,3)   (gaussian_curve_model.py:70-78)"""
    bez = (1 - t) ** 3 * cp[:, 0, :] + 3 * (1 - t) ** 2 * t * cp[:, 1, :] \
        + 3 * (1 - t) * t ** 2 * cp[:, 2, :] + t ** 3 * cp[:, 3, :]
    if bool(is_bezier.all()):
        return bez
    line = (1 - t) * cp[:, 0, :] + t * cp[:, 3, :]
    return torch.where(is_bezier.unsqueeze(0).unsqueeze(2), bez, line)


def _curve_tangent_at(cp, is_bezier, t):
    """(gaussian_curve_model.py:80-89)"""
    bez = 3 * (1 - t) ** 2 * (cp[:, 1, :] - cp[:, 0, :]) + 6 * (1 - t) * t * (cp[:, 2, :] - cp[:, 1, :]) \
        + 3 * t ** 2 * (cp[:, 3, :] - cp[:, 2, :])
    if bool(is_bezier.all()):
        return bez
    line = cp[:, 3, :] - cp[:, 0, :]
    line = line.unsqueeze(0).expand_as(bez)
    return torch.where(is_bezier.unsqueeze(0).unsqueeze(2), bez, line)


def _sqrt_positive_part(x):
    ret = torch.zeros_like(x)
    m = x > 0
    ret[m] = torch.sqrt(x[m])
    return ret


def rot_to_quat(rot):
    """(..,3,3) -> (..,4), real part first, w >= 0   (utils/general_utils.py:33-86)"""
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = torch.unbind(rot.reshape(-1, 9), dim=-1)
    q_abs = _sqrt_positive_part(torch.stack([1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22,
                                             1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22], dim=-1))
    cand = torch.stack([
        torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], dim=-1),
        torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], dim=-1),
        torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], dim=-1),
        torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], dim=-1)], dim=-2)
    flr = torch.tensor(0.1, dtype=q_abs.dtype, device=q_abs.device)
    cand = cand / (2.0 * q_abs[..., None].max(flr))
    out = cand[F.one_hot(q_abs.argmax(dim=-1), num_classes=4) > 0.5, :].reshape(-1, 4)
    return torch.where(out[..., 0:1] < 0, -out, out)


def sample_curves(curve_points, width, is_bezier, n, eps=1e-8):
    """prepare_scaling_rot (gaussian_curve_model.py:180-198) -> (xyz (P,3), rotation (P,4), scaling (P,3))."""
    dev = curve_points.device
    t = sample_t(n, dev)[:, None, None]
    xyz = _curve_points_at(curve_points, is_bezier, t)
    front = _curve_points_at(curve_points, is_bezier, t - 0.5 / n)
    dist = torch.norm(xyz - front, dim=-1)                        # (n,B)
    tangent = _curve_tangent_at(curve_points, is_bezier, t)
    B = curve_points.shape[0]
    xyz = xyz.permute(1, 0, 2).reshape(B * n, 3)                  # 'm b c -> (b m) c'
    tangent = tangent.permute(1, 0, 2).reshape(B * n, 3)
    v0 = tangent / (torch.linalg.vector_norm(tangent, dim=-1, keepdim=True) + eps)
    up = torch.tensor([[0.0, 0.0, 1.0]], device=dev, dtype=tangent.dtype)
    v1 = torch.linalg.cross(tangent, up.expand_as(tangent), dim=-1)
    v1 = v1 / torch.norm(v1)                                      # whole-tensor norm (:190)
    v2 = torch.linalg.cross(tangent, v1, dim=-1)
    v2 = v2 / torch.norm(v2)                                      # whole-tensor norm (:192)
    rot = torch.stack((v0, v1, v2), dim=1).transpose(-2, -1)
    q = rot_to_quat(rot)
    s0 = dist.permute(1, 0).reshape(B * n)
    s1 = torch.exp(width).repeat(1, n).reshape(B * n)
    return xyz, q, torch.stack((s0, s1, s1), dim=1)


def quaternion_to_matrix(q):
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(q.shape[:-1] + (3, 3))


def raster_inputs(xyz, rotation, scaling, opacity_logit, n, mask_logit, cam_center, world_view, use_mask=False,
                  mask_thr=0.01):
    """What render() hands the rasterizer (gaussian_renderer/__init__.py:57-104)."""
    rot_n = F.normalize(rotation)
    opacity = torch.sigmoid(opacity_logit.unsqueeze(1).expand(-1, n, -1).reshape(-1, 1))
    scales = scaling
    if use_mask:
        sm = torch.sigmoid(mask_logit)
        mask = ((sm > mask_thr).float() - sm).detach() + sm
        scales = scaling * mask.view(-1, 1)
        opacity = opacity * mask.view(-1, 1)
    Rm = quaternion_to_matrix(rot_n)
    dir_global = Rm[..., 0].clone()
    to_cam = cam_center - xyz
    neg = (dir_global * to_cam).sum(-1) < 0.0
    dir_global = torch.where(neg.unsqueeze(-1), -dir_global, dir_global)
    local = dir_global @ world_view[:3, :3]
    all_map = torch.cat([local, torch.ones_like(local[:, :1])], dim=1)
    colors = torch.ones(xyz.shape[0], 1, device=xyz.device)
    return xyz, opacity, scales, rot_n, colors, all_map


def curve_smoothness(rotation, n):
    """train.py:119-124 restated: dir = first column of quaternion_to_matrix(F.normalize(_rotation))
    (scene/gaussian_curve_model.py:95-97,120-122), 1 - |cosine similarity| of adjacent samples, mean."""
    q = torch.nn.functional.normalize(rotation)
    dir_global = quaternion_to_matrix(q)[..., 0].view(-1, n, 3)
    cos_sim = 1 - torch.nn.functional.cosine_similarity(dir_global[:, :-1, :], dir_global[:, 1:, :], dim=-1).abs()
    return cos_sim.mean()


def endpoint_connectivity(curve_points, dis_thr=0.05):
    """train.py:133-146 restated (exact-difference distances; torch.cdist's matmul shortcut is an fp32
    approximation of the same quantity). Returns 0 where the reference skips the term (no valid pair)."""
    start_points, end_points = curve_points[:, 0], curve_points[:, -1]
    all_points = torch.cat([start_points, end_points], dim=0)
    B = start_points.shape[0]
    mask = torch.eye(B, dtype=torch.bool, device=curve_points.device)
    mask = torch.cat([torch.cat([mask, mask], dim=1), torch.cat([mask, mask], dim=1)], dim=0)
    dist = torch.cdist(all_points, all_points, p=2, compute_mode="donot_use_mm_for_euclid_dist")
    with torch.no_grad():
        valid_mask = (dist < dis_thr) & (~mask)
    if valid_mask.any():
        return dist[valid_mask].mean()
    return curve_points.sum() * 0.0
