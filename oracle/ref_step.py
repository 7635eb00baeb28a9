"""TEST INFRASTRUCTURE ONLY (never imported by the product): runs the REFERENCE's own training step on the GPU.

The reference's Python (scene/gaussian_curve_model.py:55-198 model + sampling, gaussian_renderer/__init__.py:18-157
render(), utils/loss_utils.py:94-115 edge_aware_loss, fused_ssim/__init__.py, diff_cur_rasterization/__init__.py) is
staged unmodified by oracle/build_ref.sh into the git-ignored oracle/_ref/py/ next to the reference's CUDA extensions
recompiled for sm_100 (oracle/_ref/*.so). This script imports those files under their own module names, feeds them a
seeded synthetic scene (curve_gaussian_b200.synth: plain torch/numpy input generation), executes train.py:95-148
(render -> edge loss + fused SSIM [+ the curve-side regularisers] -> backward) and writes the image, the loss,
dL/d{_curve_points, _width, _opacity, _mask}, and the sampled Gaussians with their gradients (the boundary between the
sampling stage and the rasterizer) to an .npz. It runs as a SUBPROCESS of the parity tests and of bench.py's
`reference_gpu_step` leg, so the reference's module names (scene, utils, gaussian_renderer, ...) never meet the repo's.

Third-party modules the reference imports at module level but never uses on this path (open3d, skimage, plyfile,
seaborn, matplotlib, edge_extraction, utils.vis_utils) are stubbed; pytorch3d.transforms.quaternion_to_matrix (absent
from this image, un-pinned upstream) is restated from its published formula.

usage: python oracle/ref_step.py --spec '{"B":417,"n":12,"W":800,"H":800,...}' --out /tmp/x.npz [--time-steps K]
"""
from __future__ import annotations

import argparse
import importlib.machinery
import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")
REFPY = os.path.join(REF, "py")


class _Anything(types.ModuleType):
    """Stub module: any attribute is a harmless placeholder (only import-time names are needed)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return object


def quaternion_to_matrix(q):
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(q.shape[:-1] + (3, 3))


def _load_ext(modname, filename):
    path = os.path.join(REF, filename)
    loader = importlib.machinery.ExtensionFileLoader(modname, path)
    spec = importlib.util.spec_from_loader(modname, loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    sys.modules[modname] = mod
    return mod


def import_reference():
    """-> (GaussianCurveModel, render, edge_aware_loss, fused_ssim) of the reference, unmodified."""
    if not os.path.isdir(REFPY):
        raise RuntimeError("oracle/_ref/py missing: run oracle/build_ref.sh where /root/reference exists")
    sys.path.insert(0, REFPY)
    _load_ext("diff_cur_rasterization._C", "diff_cur_rasterization_C.so")
    _load_ext("fused_ssim_cuda", "fused_ssim_cuda.so")
    sys.modules.setdefault("simple_knn", types.ModuleType("simple_knn"))
    _load_ext("simple_knn._C", "simple_knn_C.so")
    p3 = _Anything("pytorch3d.transforms")
    p3.quaternion_to_matrix = quaternion_to_matrix
    sys.modules["pytorch3d"] = _Anything("pytorch3d")
    sys.modules["pytorch3d.transforms"] = p3
    for _ in range(80):
        try:
            from scene.gaussian_curve_model import GaussianCurveModel
            from gaussian_renderer import render
            from utils.loss_utils import edge_aware_loss
            from fused_ssim import fused_ssim
            return GaussianCurveModel, render, edge_aware_loss, fused_ssim
        except ModuleNotFoundError as e:
            parts = e.name.split(".")
            for i in range(1, len(parts) + 1):
                n = ".".join(parts[:i])
                if n not in sys.modules:
                    sys.modules[n] = _Anything(n)
    raise RuntimeError("could not import the reference")


class Pipe:   # arguments/__init__.py:66-75 defaults
    convert_SHs_python = False
    compute_cov3D_python = False
    debug = False
    antialiasing = False
    render_geo = True


def build_model(GaussianCurveModel, spec, dev):
    sys.path.insert(0, ROOT)
    from curve_gaussian_b200 import synth
    B, n = spec["B"], spec["n"]
    cp, width, opl, isb = synth.random_curves(B, seed=spec.get("seed", 0), line_fraction=spec.get("line_fraction", 0.0))
    if "spread" in spec:
        cp = (cp - 0.5) * spec["spread"] + 0.5
    width = width + spec.get("width_shift", 0.0)
    m = GaussianCurveModel(0, n_gaussians=n)
    P = lambda t: torch.nn.Parameter(t.to(dev).contiguous().requires_grad_(True))
    m._curve_points = P(cp)
    m._width = P(width)
    m._opacity = P(opl)
    if spec.get("use_mask", False):
        g = torch.Generator().manual_seed(spec.get("seed", 0) + 5)
        m._mask = P(torch.randn(B, n, 1, generator=g) * 3)
    else:
        m._mask = P(torch.ones(B, n, 1))
    m.is_bezier = isb.to(dev)
    m._features_dc = torch.zeros(B, n, 1, 1, device=dev)
    m._features_rest = torch.zeros(B, n, 0, 1, device=dev)
    cams = synth.random_cameras(spec.get("views", 1), spec["W"], spec["H"], seed=spec.get("cam_seed", 2))
    return m, [c.to(dev) for c in cams], (cp, width, opl, isb)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--spec", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--time-steps", type=int, default=0)
    ap.add_argument("--repeats", type=int, default=3, help="independent fwd+bwd runs saved (the reference's own fp32-atomic noise)")
    a = ap.parse_args()
    spec = json.loads(a.spec)
    dev = torch.device("cuda:0")
    GaussianCurveModel, render, edge_aware_loss, fused_ssim = import_reference()
    import torch.nn.functional as F
    from einops import rearrange
    torch.manual_seed(0)
    m, cams, _ = build_model(GaussianCurveModel, spec, dev)
    cam = cams[0]
    bg = torch.zeros(3, device=dev)
    use_mask = bool(spec.get("use_mask", False))
    regs = bool(spec.get("regs", False))
    lambda_mse, lambda_dssim = spec.get("lambda_mse", 10.0), spec.get("lambda_dssim", 0.1)   # arguments/__init__.py:94,98

    # ground-truth edge map: the reference's own render of the same curves with perturbed control points, binarised
    with torch.no_grad():
        keep = m._curve_points.data.clone()
        g = torch.Generator().manual_seed(spec.get("seed", 0) + 77)
        m._curve_points.data += (torch.randn(keep.shape, generator=g) * spec.get("gt_jitter", 0.01)).to(dev)
        m.prepare_scaling_rot()
        gt = (render(cam, m, Pipe(), bg, use_mask=use_mask)["render"] > 0.3).float()
        m._curve_points.data.copy_(keep)

    def step():
        for p in (m._curve_points, m._width, m._opacity, m._mask):
            p.grad = None
        m.prepare_scaling_rot()                                              # train.py:86 (scene/gaussian_curve_model.py:180-198)
        for t in (m._xyz, m._rotation, m._scaling):                          # (keeps dL/d(sampled Gaussians) for the stage-wise parity tests)
            t.retain_grad()
        pkg = render(cam, m, Pipe(), bg, use_mask=use_mask, mask_thr=0.01)   # train.py:95-97
        image, vis = pkg["render"], pkg["visibility_filter"]
        Ll1 = edge_aware_loss(image, gt[:1])                                 # train.py:101
        ssim_value = fused_ssim(image.unsqueeze(0), gt[:1].unsqueeze(0))     # train.py:103
        loss = lambda_mse * ((1.0 - lambda_dssim) * Ll1 + lambda_dssim * (1.0 - ssim_value))   # train.py:107
        if regs:
            if use_mask:
                loss = loss + 0.0005 * torch.mean(torch.sigmoid(m._mask))   # train.py:110-111
            if vis.sum() > 0:                                                # train.py:114-117
                opacity = m.get_opacity[vis]
                loss = loss + 0.01 * torch.log(1 + opacity ** 2 / 0.5).mean()
                rotation_mat = m.get_rotation_matrix                         # train.py:119-124
                dir_global = rearrange(rotation_mat[..., 0], '(b m) c -> b m c', m=m.n_gaussians)
                cos_sim = 1 - F.cosine_similarity(dir_global[:, :-1, :], dir_global[:, 1:, :], dim=-1).abs()
                loss = loss + 0.1 * cos_sim.mean()
            mask = m.get_curve_width >= 0.005                                # train.py:126-131
            if mask.any():
                loss = loss + 0.01 * (m.get_curve_width[mask] - 0.005).mean()
        loss.backward()
        return pkg, loss

    out = {"gt": gt.cpu().numpy()}
    for r in range(max(1, a.repeats)):
        pkg, loss = step()
        torch.cuda.synchronize()
        tag = "" if r == 0 else f"_run{r}"
        out["g_curve_points" + tag] = m._curve_points.grad.cpu().numpy()
        out["g_width" + tag] = m._width.grad.cpu().numpy()
        out["g_opacity" + tag] = m._opacity.grad.cpu().numpy()
        out["g_mask" + tag] = (m._mask.grad if m._mask.grad is not None else torch.zeros_like(m._mask)).cpu().numpy()
        out["loss" + tag] = np.float64(loss.item())
        out["g_xyz" + tag] = m._xyz.grad.cpu().numpy()
        out["g_rotation" + tag] = m._rotation.grad.cpu().numpy()
        out["g_scaling" + tag] = m._scaling.grad.cpu().numpy()
        if r == 0:
            out["image"] = pkg["render"].detach().cpu().numpy()
            out["xyz"] = m._xyz.detach().cpu().numpy()
            out["rotation"] = m._rotation.detach().cpu().numpy()
            out["scaling"] = m._scaling.detach().cpu().numpy()
            out["radii"] = pkg["radii"].cpu().numpy()
            out["rend_dir"] = pkg["rend_dir"].detach().cpu().numpy()
            out["rend_alpha"] = pkg["rend_alpha"].detach().cpu().numpy()
            out["depth"] = pkg["depth"].detach().cpu().numpy()
    if a.time_steps > 0:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.time_steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        out["ms_per_step"] = np.float64(e0.elapsed_time(e1) / a.time_steps)
    out["repeats"] = np.int64(max(1, a.repeats))
    np.savez(a.out, **out)
    print(json.dumps({"ok": True, "loss": float(out["loss"]), "ms_per_step": float(out.get("ms_per_step", 0.0))}))


if __name__ == "__main__":
    main()
