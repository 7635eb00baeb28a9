"""Host-side mirror of the reference GaussianCurveModel (scene/gaussian_curve_model.py): same attribute
names, properties and `prepare_scaling_rot()`, with the sampling math running as the fused CUDA op in
sampling.py (:54-198, the hot path). The training-time surgery on the curve set and the optimizer
bookkeeping (:200-459) come from topology.CurveTopology; checkpoints and the on-disk formats (ply,
parametric_edges.json) from curve_io. The open3d mesh dump (:643-711) is not carried over.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from . import curve_io, sampling
from .knn import distCUDA2
from .topology import CurveTopology


def quaternion_to_matrix(quaternions: torch.Tensor) -> torch.Tensor:
    """Real-first quaternion -> rotation matrix (pytorch3d.transforms formula; the reference
    imports it at scene/gaussian_curve_model.py:6 and calls it at :97)."""
    r, i, j, k = torch.unbind(quaternions, -1)
    two_s = 2.0 / (quaternions * quaternions).sum(-1)
    o = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(quaternions.shape[:-1] + (3, 3))


def initialize_bezier_curves(points, bound, n_control_points=4):
    """One vertical cubic Bezier per point (scene/gaussian_curve_model.py:27-51)."""
    assert n_control_points == 4
    direction = torch.cat([torch.zeros_like(bound), bound, torch.zeros_like(bound)], dim=1)
    return torch.stack([points - direction, points - 0.5 * direction, points + 0.5 * direction, points + direction],
                       dim=1)


class GaussianCurveModel(CurveTopology):
    def __init__(self, sh_degree=0, n_gaussians=12, optimizer_type="default", device="cuda"):
        self.active_sh_degree = 0
        self.max_sh_degree = sh_degree
        self.optimizer_type = optimizer_type
        self.n_gaussians = n_gaussians
        t = sampling.sample_t(n_gaussians, device)
        self.sample_t = t[:, None, None]
        self._xyz = torch.empty(0)
        self._scaling = torch.empty(0)
        self._rotation = torch.empty(0)
        self._opacity = torch.empty(0)
        self._width = torch.empty(0)
        self._mask = torch.empty(0)
        self._curve_points = torch.empty(0)
        self.is_bezier = torch.empty(0)
        self.max_radii2D = torch.empty(0)
        self._features_dc = torch.empty(0)
        self._features_rest = torch.empty(0)
        self.xyz_gradient_accum = torch.empty(0)
        self.denom = torch.empty(0)
        self.optimizer = None
        self.spatial_lr_scale = 0
        # activations of the base class (scene/gaussian_model.py:38-53)
        self.scaling_activation = torch.exp
        self.scaling_inverse_activation = torch.log
        self.opacity_activation = torch.sigmoid
        self.inverse_opacity_activation = lambda x: torch.log(x / (1 - x))
        self.rotation_activation = F.normalize

    # ---- construction ---------------------------------------------------
    def create_from_curves(self, curve_points, width, opacity_logit, is_bezier=None, mask=None):
        dev = self.sample_t.device
        B = curve_points.shape[0]
        own = lambda t, *shape: nn.Parameter(t.detach().to(dev, torch.float32).reshape(*shape).clone().requires_grad_(True))
        self._curve_points = own(curve_points, B, 4, 3)
        self._width = own(width, B, 1)
        self._opacity = own(opacity_logit, B, 1)
        if mask is None:
            mask = torch.ones((B, self.n_gaussians, 1), device=dev)
        self._mask = own(mask, B, self.n_gaussians, 1)
        self.is_bezier = (torch.ones(B, dtype=torch.bool, device=dev) if is_bezier is None
                          else is_bezier.to(dev).bool())
        self.max_radii2D = torch.zeros(B * self.n_gaussians, device=dev)
        # per-sample SH feature slots of the reference (:159-166); render() never reads them (colours are a
        # constant 1), they exist so that the optimizer groups and checkpoints keep the reference's layout
        k = (self.max_sh_degree + 1) ** 2 - 1
        self._features_dc = nn.Parameter(torch.zeros((B, self.n_gaussians, 1, 1), device=dev).requires_grad_(True))
        self._features_rest = nn.Parameter(torch.zeros((B, self.n_gaussians, k, 1), device=dev).requires_grad_(True))
        self.prepare_scaling_rot()
        return self

    def create_from_pcd(self, pcd, cam_infos, spatial_lr_scale: float, init_size: float = 0.5, n_control_points: int = 4):
        """scene/gaussian_curve_model.py:142-178: one vertical Bezier per input point, sized by the distance to its
        3 nearest neighbours; per-image exposure slots as the reference keeps them (`cam_infos`: objects with
        `.image_name`, or an int count)."""
        import numpy as np
        assert n_control_points == 4
        self.spatial_lr_scale = spatial_lr_scale
        dev = self.sample_t.device
        self.create_from_points(torch.as_tensor(np.asarray(pcd.points)).float(), init_size)
        colors = getattr(pcd, "colors", None)
        if colors is not None and len(colors) == self._curve_points.shape[0]:
            # RGB2SH of the first colour channel, repeated over the curve's samples (:158-166): only read back by the
            # PLY / checkpoint writers (render() draws every Gaussian with colour 1)
            c0 = torch.as_tensor(np.asarray(colors))[:, 0:1].float().to(dev)
            dc = ((c0 - 0.5) / 0.28209479177387814)[:, None, :, None].repeat(1, self.n_gaussians, 1, 1)
            self._features_dc = nn.Parameter(dc.contiguous().requires_grad_(True))
        names = [getattr(c, "image_name", str(i)) for i, c in enumerate(cam_infos)] if not isinstance(cam_infos, int) \
            else [str(i) for i in range(cam_infos)]
        self.exposure_mapping = {name: i for i, name in enumerate(names)}
        self.pretrained_exposures = None
        self._exposure = nn.Parameter(torch.eye(3, 4, device=dev)[None].repeat(max(len(names), 1), 1, 1).requires_grad_(True))
        return self

    def create_from_points(self, points, init_size=0.5):
        """The geometric part of create_from_pcd (scene/gaussian_curve_model.py:142-178)."""
        dev = self.sample_t.device
        pts = points.to(dev).float()
        dist2 = torch.clamp_min(distCUDA2(pts), 0.0000001)
        self.dist = torch.sqrt(dist2).mean()
        bound = init_size * torch.sqrt(dist2).unsqueeze(1)
        cp = initialize_bezier_curves(pts, bound)
        n = pts.shape[0]
        opac = self.inverse_opacity_activation(0.6 * torch.ones((n, 1), dtype=torch.float, device=dev))
        widths = self.scaling_inverse_activation(5e-3 * torch.ones((n, 1), dtype=torch.float, device=dev))
        return self.create_from_curves(cp, widths, opac)

    # ---- the hot path -----------------------------------------------------
    def prepare_scaling_rot(self, eps=1e-8):
        # drop the previous iteration's autograd graph BEFORE building the new one: while the old sampled
        # tensors are alive they keep the parameters' AccumulateGrad nodes (and the stream those were created
        # on) alive, which ties a step captured into a CUDA graph to the stream of an earlier eager step
        self._xyz = self._rotation = self._scaling = None
        xyz, rot, scaling = sampling.sample_curves(self._curve_points, self._width, self.is_bezier,
                                                   self.sample_t.view(-1))
        self._xyz, self._rotation, self._scaling = xyz, rot, scaling

    # ---- accessors (same names as the reference) -------------------------
    @property
    def get_curve_points(self):
        return self._curve_points

    @property
    def get_scaling(self):
        return self._scaling

    @property
    def get_rotation_matrix(self):
        return quaternion_to_matrix(self.get_rotation)

    def get_main_axis(self, view_cam):
        dir_global = self.get_rotation_matrix[..., 0]
        to_cam = view_cam.camera_center - self._xyz
        neg_mask = (dir_global * to_cam).sum(-1) < 0.0
        return torch.where(neg_mask.unsqueeze(-1), -dir_global, dir_global)

    @property
    def get_opacity(self):
        return self.opacity_activation(self._opacity.unsqueeze(1).expand(-1, self.n_gaussians, -1).reshape(-1, 1))

    @property
    def get_curve_opacity(self):
        return self.opacity_activation(self._opacity)

    @property
    def get_curve_width(self):
        return self.scaling_activation(self._width)

    @property
    def get_rotation(self):
        return self.rotation_activation(self._rotation)

    @property
    def get_xyz(self):
        return self._xyz

    def parameters(self):
        return [self._curve_points, self._width, self._opacity, self._mask]

    # ---- the remaining calls train.py makes on the model ---------------------
    def oneupSHdegree(self):
        if self.active_sh_degree < self.max_sh_degree:
            self.active_sh_degree += 1

    def get_exposure_from_name(self, image_name):
        if getattr(self, "pretrained_exposures", None) is None:
            return self._exposure[self.exposure_mapping[image_name]]
        return self.pretrained_exposures[image_name]

    @torch.no_grad()
    def draw_curve(self, path, step, num_sample=200):
        """`curve_step{step}.ply`: num_sample points per curve (gaussian_curve_model.py:713-727, without colours)."""
        cp = self._curve_points
        t = torch.linspace(0, 1, num_sample, device=cp.device)[None, :, None]
        p0, p1, p2, p3 = (cp[:, i][:, None, :] for i in range(4))
        bez = (1 - t) ** 3 * p0 + 3 * (1 - t) ** 2 * t * p1 + 3 * (1 - t) * t ** 2 * p2 + t ** 3 * p3
        pts = torch.where(self.is_bezier[:, None, None], bez, (1 - t) * p0 + t * p3)
        curve_io.write_ascii_points(f"{path}/curve_step{step}.ply", pts.reshape(-1, 3).cpu().numpy())

    def draw_ellipsoids(self, path, step, radius=1.2):
        """The reference dumps one open3d sphere mesh per Gaussian here (a debug visualisation); not carried over."""

    def load_ply(self, path, use_train_test_exp=False):
        raise NotImplementedError("a per-Gaussian point_cloud.ply does not determine the curve parameters "
                                  "(the reference's loader cannot rebuild them either); use load_curves() / restore()")

    # ---- checkpoints and on-disk formats (curve_io) -------------------------
    def capture(self):
        return curve_io.capture(self)

    def restore(self, model_args, training_args=None):
        return curve_io.restore(self, model_args, training_args)

    def save_ply(self, path):
        return curve_io.save_gaussians_ply(self, path)

    def save_curves(self, path):
        return curve_io.save_curves(self, path)

    def load_curves(self, path):
        return curve_io.load_curves(self, path)
