"""Generates tests/golden/sampling_*.npz by running the REFERENCE's own Python
(scene/gaussian_curve_model.py prepare_scaling_rot + utils/general_utils.py
rot_to_quat_batch, and the render()-side activations) on CPU in the build
container. Third-party modules the reference imports but this image lacks are
stubbed; pytorch3d.transforms.quaternion_to_matrix is restated (formula in
SURVEY.md 8c, un-pinned upstream). Run:  python tests/golden/make_sampling_golden.py
Needs /root/reference; the .npz outputs are committed, this script only documents them.
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("REF", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from curve_gaussian_b200 import synth  # noqa: E402


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def quaternion_to_matrix(q):
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(q.shape[:-1] + (3, 3))


class _Anything(types.ModuleType):
    """Stub module: any attribute is a harmless placeholder (only import-time names are needed)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return object


def import_reference_model():
    sys.path.insert(0, REF)
    for _ in range(60):
        try:
            from scene.gaussian_curve_model import GaussianCurveModel
            return GaussianCurveModel
        except ModuleNotFoundError as e:
            parts = e.name.split(".")
            for i in range(1, len(parts) + 1):
                n = ".".join(parts[:i])
                if n not in sys.modules:
                    sys.modules[n] = _Anything(n)
            if e.name.startswith("pytorch3d"):
                sys.modules.setdefault("pytorch3d.transforms", _Anything("pytorch3d.transforms"))
            if "pytorch3d.transforms" in sys.modules:
                sys.modules["pytorch3d.transforms"].quaternion_to_matrix = quaternion_to_matrix
    raise RuntimeError("could not import the reference model")


def run_case(GaussianCurveModel, name, B, n, line_fraction, seed, use_mask):
    cp, width, opl, isb = synth.random_curves(B, seed=seed, line_fraction=line_fraction)
    m = GaussianCurveModel.__new__(GaussianCurveModel)   # __init__ hard-codes device='cuda' (:58-59)
    m.n_gaussians = n
    m.sample_t = torch.linspace(0.5 / n, 1 - 0.5 / n, n)[:, None, None]
    m.opacity_activation = torch.sigmoid
    m.scaling_activation = torch.exp
    m.rotation_activation = torch.nn.functional.normalize
    m._curve_points = cp.clone().requires_grad_(True)
    m._width = width.clone().requires_grad_(True)
    m._opacity = opl.clone().requires_grad_(True)
    g = torch.Generator().manual_seed(seed + 5)
    m._mask = (torch.randn(B, n, 1, generator=g) * 3).requires_grad_(True)
    m.is_bezier = isb
    m.prepare_scaling_rot()
    cam = synth.random_cameras(1, 64, 48, seed=seed)[0]
    xyz, rot, scal = m._xyz, m._rotation, m._scaling
    opacity = m.get_opacity
    rot_n = m.get_rotation
    scales = m.get_scaling
    if use_mask:   # gaussian_renderer/__init__.py:72-76
        mask = ((torch.sigmoid(m._mask) > 0.01).float() - torch.sigmoid(m._mask)).detach() + torch.sigmoid(m._mask)
        scales = m.get_scaling * mask.view(-1, 1)
        opacity = m.get_opacity * mask.view(-1, 1)
    view_cam = types.SimpleNamespace(camera_center=cam.camera_center)
    axis = m.get_main_axis(view_cam)                       # gaussian_curve_model.py:99-105
    local = axis @ cam.world_view_transform[:3, :3]        # gaussian_renderer/__init__.py:98-99
    gg = torch.Generator().manual_seed(seed + 9)
    outs = (xyz, rot, scal, opacity, rot_n, scales, local)
    w = [torch.randn(t.shape, generator=gg) for t in outs]
    loss = sum((a * b).sum() for a, b in zip(outs, w))
    loss.backward()
    out = dict(curve_points=cp, width=width, opacity_logit=opl, mask_logit=m._mask.detach(), is_bezier=isb,
               n=np.int64(n), use_mask=np.bool_(use_mask), cam_center=cam.camera_center,
               world_view=cam.world_view_transform,
               xyz=xyz, rotation=rot, scaling=scal, opacity=opacity, rot_normalized=rot_n, scales_masked=scales,
               local_axis=local, w_xyz=w[0], w_rot=w[1], w_scal=w[2], w_opacity=w[3], w_rot_n=w[4], w_scales=w[5],
               w_local=w[6], g_curve_points=m._curve_points.grad, g_width=m._width.grad, g_opacity=m._opacity.grad,
               g_mask=m._mask.grad if m._mask.grad is not None else torch.zeros_like(m._mask))
    out = {k: (v.detach().numpy() if torch.is_tensor(v) else v) for k, v in out.items()}
    np.savez_compressed(os.path.join(HERE, f"sampling_{name}.npz"), **out)
    print(name, "P =", B * n, "saved")


if __name__ == "__main__":
    torch.manual_seed(0)
    G = import_reference_model()
    run_case(G, "bezier1x32", 1, 32, 0.0, 0, False)        # BASELINE config[0] shape
    run_case(G, "bezier40x12", 40, 12, 0.0, 1, False)      # default n_gaussians
    run_case(G, "mixed25x16_mask", 25, 16, 0.3, 2, True)   # lines + mask straight-through
