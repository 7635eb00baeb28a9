"""Checkpoints and on-disk formats of the curve model.

* `capture` / `restore`: the reference's checkpoint tuple (scene/gaussian_model.py:74-106) plus the curve
  parameters the reference forgets to save (`_curve_points/_width/_mask/is_bezier` are absent from its tuple,
  so a restored reference model cannot re-sample its Gaussians).
* `save_gaussians_ply`: the per-Gaussian `point_cloud.ply` (scene/gaussian_model.py:383-400; same attribute
  names and order, binary little-endian float32), written without the `plyfile` dependency.
* `extract_curves`: `parametric_edges.json` + `edge_points.ply` as train.py:250-293 writes them for the
  evaluation scripts, including the end-point merging (`merge_endpoints`, edge_extraction/merging.py:10-55); the
  visibility filtering option (`opt.visible_checking`, off by default, needs the dataset's edge maps) is not part
  of this port.
"""
from __future__ import annotations

import json
import os
from typing import Dict, Tuple

import numpy as np
import torch
from torch import nn

CURVE_KEYS = ("curve_points", "width", "mask", "is_bezier", "n_gaussians")


# ---- checkpoint tuple ------------------------------------------------------------------------------------------
def capture(model) -> tuple:
    """(the reference's 12 entries ..., curve dict). Index 12 is ours."""
    opt = model.optimizer.state_dict() if getattr(model, "optimizer", None) is not None else None
    if getattr(model, "_curve_points", None) is not None and model._curve_points.is_cuda and hasattr(model, "prepare_scaling_rot"):
        # a graph-mode loop re-samples inside its captured step only: make the sampled tensors saved below belong to
        # the curve parameters saved next to them
        with torch.no_grad():
            model.prepare_scaling_rot()
    params = ("_curve_points", "_width", "_opacity", "_mask", "_features_dc", "_features_rest")
    curves = {"curve_points": model._curve_points, "width": model._width, "mask": model._mask,
              "is_bezier": model.is_bezier, "n_gaussians": model.n_gaussians,
              # which parameters were trainable (fix_opacity() freezes _opacity), and the per-image exposure state
              "requires_grad": {k: bool(getattr(model, k).requires_grad) for k in params if hasattr(model, k)},
              "exposure": getattr(model, "_exposure", None),
              "exposure_mapping": dict(getattr(model, "exposure_mapping", {}) or {})}
    return (model.active_sh_degree, model._xyz, model._features_dc, model._features_rest, model._scaling,
            model._rotation, model._opacity, model.max_radii2D, model.xyz_gradient_accum, model.denom, opt,
            model.spatial_lr_scale, curves)


def restore(model, model_args, training_args=None) -> None:
    if len(model_args) < 13:
        raise ValueError("this checkpoint has the reference's 12 entries only: it does not contain the curve "
                         "parameters (_curve_points/_width/_mask/is_bezier), so the curve model cannot be rebuilt")
    (model.active_sh_degree, _xyz, f_dc, f_rest, _scaling, _rotation, opacity, max_radii2D, accum, denom, opt_dict,
     model.spatial_lr_scale, curves) = model_args[:13]
    if int(curves["n_gaussians"]) != int(model.n_gaussians):
        raise ValueError(f"checkpoint samples {curves['n_gaussians']} Gaussians per curve, the model {model.n_gaussians}")
    dev = model.sample_t.device
    rg = curves.get("requires_grad", {})
    P = lambda t, k: nn.Parameter(t.detach().to(dev).clone().requires_grad_(bool(rg.get(k, True))))
    model._curve_points, model._width = P(curves["curve_points"], "_curve_points"), P(curves["width"], "_width")
    model._mask = P(curves["mask"], "_mask")
    model._opacity, model._features_dc = P(opacity, "_opacity"), P(f_dc, "_features_dc")
    model._features_rest = P(f_rest, "_features_rest")
    if curves.get("exposure") is not None:
        model._exposure = nn.Parameter(curves["exposure"].detach().to(dev).clone().requires_grad_(True))
        model.exposure_mapping = dict(curves.get("exposure_mapping", {}))
    model.is_bezier = curves["is_bezier"].to(dev).bool().clone()
    model.max_radii2D = max_radii2D.to(dev).clone()
    model.prepare_scaling_rot()          # _xyz/_rotation/_scaling are functions of the curve parameters
    if training_args is not None:
        model.training_setup(training_args)
        if opt_dict is not None:
            model.optimizer.load_state_dict(opt_dict)
    if accum is not None and accum.numel():
        model.xyz_gradient_accum, model.denom = accum.to(dev).clone(), denom.to(dev).clone()


def save_curves(model, path: str) -> None:
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save({"curve_points": model._curve_points.detach().cpu(), "width": model._width.detach().cpu(),
                "opacity": model._opacity.detach().cpu(), "mask": model._mask.detach().cpu(),
                "is_bezier": model.is_bezier.cpu(), "n_gaussians": model.n_gaussians}, path)


def load_curves(model, path: str):
    d = torch.load(path, map_location="cpu")
    if int(d["n_gaussians"]) != int(model.n_gaussians):
        raise ValueError(f"file samples {d['n_gaussians']} Gaussians per curve, the model {model.n_gaussians}")
    return model.create_from_curves(d["curve_points"], d["width"], d["opacity"], d["is_bezier"], d["mask"])


# ---- PLY -------------------------------------------------------------------------------------------------------
def ply_attributes(n_dc: int, n_rest: int) -> list:
    """construct_list_of_attributes (scene/gaussian_model.py:267-280)."""
    names = ["x", "y", "z", "nx", "ny", "nz"]
    names += [f"f_dc_{i}" for i in range(n_dc)] + [f"f_rest_{i}" for i in range(n_rest)] + ["opacity"]
    return names + [f"scale_{i}" for i in range(3)] + [f"rot_{i}" for i in range(4)]


def write_ply(path: str, names, table: np.ndarray) -> None:
    """binary_little_endian PLY with one float32 property per column of `table` (N, len(names))."""
    table = np.ascontiguousarray(table, dtype="<f4")
    assert table.ndim == 2 and table.shape[1] == len(names)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    head = "ply\nformat binary_little_endian 1.0\n" + f"element vertex {table.shape[0]}\n" + \
        "".join(f"property float {n}\n" for n in names) + "end_header\n"
    with open(path, "wb") as f:
        f.write(head.encode("ascii"))
        f.write(table.tobytes())


def read_ply(path: str) -> Dict[str, np.ndarray]:
    """Reader for the files write_ply / write_ascii_points produce (float / double vertex properties)."""
    with open(path, "rb") as f:
        assert f.readline().strip() == b"ply"
        fmt, count, props = None, 0, []
        while True:
            line = f.readline().decode("ascii").strip()
            if line == "end_header":
                break
            tok = line.split()
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[:2] == ["element", "vertex"]:
                count = int(tok[2])
            elif tok[0] == "property":
                props.append((tok[2], {"float": "<f4", "double": "<f8"}[tok[1]]))
        if fmt == "binary_little_endian":
            data = np.frombuffer(f.read(), dtype=np.dtype(props), count=count)
            return {n: np.array(data[n]) for n, _ in props}
        rows = np.loadtxt(f, ndmin=2) if count else np.zeros((0, len(props)))
        return {n: rows[:, i] for i, (n, _) in enumerate(props)}


def save_gaussians_ply(model, path: str) -> None:
    n = model.n_gaussians
    xyz = model._xyz.detach()
    P = xyz.shape[0]
    opac = model.inverse_opacity_activation(model.get_opacity).detach()
    f_dc = model._features_dc.detach().reshape(P, -1)
    f_rest = model._features_rest.detach().reshape(P, -1)
    table = torch.cat([xyz, torch.zeros_like(xyz), f_dc, f_rest, opac.view(P, 1), model._scaling.detach(),
                       model._rotation.detach()], dim=1).cpu().numpy()
    write_ply(path, ply_attributes(f_dc.shape[1], f_rest.shape[1]), table)


def write_ascii_points(path: str, pts: np.ndarray) -> None:
    """ASCII point cloud with double x/y/z, the layout open3d's write_point_cloud(write_ascii=True) emits."""
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w") as f:
        f.write("ply\nformat ascii 1.0\ncomment curve-gaussian_b200 edge points\n"
                f"element vertex {len(pts)}\nproperty double x\nproperty double y\nproperty double z\nend_header\n")
        for p in np.asarray(pts, dtype=np.float64):
            f.write(f"{p[0]:.10f} {p[1]:.10f} {p[2]:.10f}\n")


# ---- parametric edges --------------------------------------------------------------------------------------------
def bezier_lengths(ctrl: np.ndarray, num_samples: int = 100) -> np.ndarray:
    """Arc length of cubic Beziers (K,4,3): composite Simpson over num_samples sub-intervals with num_samples
    panels each, the quadrature of edge_extraction/extract_uitl.py:291-336, evaluated for all curves at once."""
    ctrl = np.asarray(ctrl, dtype=np.float64).reshape(-1, 4, 3)
    m = num_samples
    h = 1.0 / (m * m)
    i = np.arange(m + 1)
    w = np.where((i == 0) | (i == m), 1.0, np.where(i % 2 == 1, 4.0, 2.0))
    t = (np.arange(m)[:, None] / m + i[None, :] * h).reshape(-1)              # (m*(m+1),)
    d = ctrl[:, 1:] - ctrl[:, :-1]                                            # (K,3,3)
    basis = np.stack([3 * (1 - t) ** 2, 6 * (1 - t) * t, 3 * t ** 2], axis=0)  # (3,T)
    speed = np.linalg.norm(np.einsum("jt,kjc->ktc", basis, d), axis=-1)        # (K,T)
    return (speed.reshape(len(ctrl), m, m + 1) * w).sum(axis=(1, 2)) * h / 3


def sample_edges(edge_dict: dict, sample_resolution: float = 0.005) -> np.ndarray:
    """Points every ~sample_resolution along each curve / line (extract_para_edge.py:107-129)."""
    curves = np.asarray(edge_dict["curves_ctl_pts"], dtype=np.float64).reshape(-1, 4, 3)
    lines = np.asarray(edge_dict["lines_end_pts"], dtype=np.float64).reshape(-1, 2, 3)
    out = []
    if len(curves):
        coeff = np.array([[-1, 3, -3, 1], [3, -6, 3, 0], [-3, 3, 0, 0], [1, 0, 0, 0]], dtype=np.float64)
        for c, length in zip(curves, bezier_lengths(curves)):
            t = np.linspace(0, 1, int(length // sample_resolution))
            out.append(np.stack([t ** 3, t ** 2, t, np.ones_like(t)], axis=1) @ coeff @ c)
    for a, b in lines:
        t = np.linspace(0, 1, int(np.linalg.norm(a - b) // sample_resolution))
        out.append(np.outer(t, b - a) + a)
    return np.concatenate(out, axis=0).astype(np.float32) if out else np.zeros((0, 3), np.float32)


def merge_endpoints(lines: np.ndarray, curves: np.ndarray, distance_threshold: float):
    """Snap end points that lie within distance_threshold of each other (transitively) to their common mean
    (edge_extraction/merging.py:10-55). lines: (Nl,6) end-point pairs, curves: (Nc,12) control polygons; only the
    first and last control point of a curve move. Returns the two arrays with merged end points."""
    lines = np.asarray(lines, dtype=np.float64).reshape(-1, 6)
    curves = np.asarray(curves, dtype=np.float64).reshape(-1, 12)
    pts = np.concatenate([lines.reshape(-1, 3), curves[:, [0, 1, 2, 9, 10, 11]].reshape(-1, 3)], axis=0)
    n = len(pts)
    if n == 0:
        return lines, curves
    adj = np.linalg.norm(pts[:, None, :] - pts[None, :, :], axis=-1) <= distance_threshold
    labels = np.arange(n)
    while True:                                   # min-label propagation = connected components
        nxt = np.where(adj, labels[None, :], n).min(axis=1)
        if np.array_equal(nxt, labels):
            break
        labels = nxt
    sums = np.zeros((n, 3))
    np.add.at(sums, labels, pts)
    counts = np.bincount(labels, minlength=n)[:, None]
    merged = np.where(counts[labels] > 1, sums[labels] / np.maximum(counts[labels], 1), pts)
    out_lines = merged[: 2 * len(lines)].reshape(-1, 6)
    ends = merged[2 * len(lines):].reshape(-1, 6)
    out_curves = curves.copy()
    out_curves[:, :3], out_curves[:, 9:] = ends[:, :3], ends[:, 3:]
    return out_lines, out_curves


def edge_dict(model, merge_endpoints_flag: bool = False, distance_threshold: float = 0.015) -> dict:
    """{'curves_ctl_pts': (Nb,4,3) lists, 'lines_end_pts': (Nl,6) lists} (train.py:252-271, extract_para_edge.py:83-100)."""
    cp = model.get_curve_points.detach()
    isb = model.is_bezier
    curves = cp[isb].cpu().double().numpy().reshape(-1, 12)
    lines = cp[~isb][:, [0, -1], :].reshape(-1, 6).cpu().double().numpy()
    if merge_endpoints_flag:
        lines, curves = merge_endpoints(lines, curves, distance_threshold)
    return {"curves_ctl_pts": curves.reshape(-1, 4, 3).tolist(), "lines_end_pts": lines.tolist()}


def extract_curves(model, model_path: str, sample_resolution: float = 0.005, merge_endpoints_flag: bool = True,
                   distance_threshold: float = 0.015) -> Tuple[np.ndarray, dict]:
    """Write `parametric_edges.json` and `edge_points.ply` under model_path (train.py:250-293; end points closer
    than distance_threshold are merged first, `opt.merge_endpoints_flag`, on by default as in the reference)."""
    d = edge_dict(model, merge_endpoints_flag, distance_threshold)
    pts = sample_edges(d, sample_resolution)
    write_ascii_points(os.path.join(model_path, "edge_points.ply"), pts)
    with open(os.path.join(model_path, "parametric_edges.json"), "w") as f:
        json.dump(d, f)
    return pts, d
