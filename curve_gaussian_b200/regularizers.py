"""Curve-side regularisers of the training step as fused CUDA ops.

`curve_smoothness(model._rotation, n)` equals train.py:119-124,

    rotation_mat = gaussians.get_rotation_matrix
    dir_global = rearrange(rotation_mat[..., 0], '(b m) c -> b m c', m=n)
    curve_smo = (1 - F.cosine_similarity(dir_global[:, :-1], dir_global[:, 1:], dim=-1).abs()).mean()

and `endpoint_connectivity(model._curve_points)` equals train.py:133-146 (mean distance of the endpoint pairs
closer than 0.05, a curve's own two endpoints excluded; 0 when no pair qualifies, where the reference skips the
term) without torch.cdist's dense (2B)^2 matrix. Both return device scalars and never synchronise the host.
"""
from __future__ import annotations

import torch

from . import _lib


class _CurveSmooth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rotation, n):
        lib = _lib.load()
        if not rotation.is_cuda:
            raise _lib.CurveGSError("curve_smoothness needs CUDA tensors; there is no CPU path")
        if rotation.ndim != 2 or rotation.shape[1] != 4 or rotation.shape[0] % n:
            raise _lib.CurveGSError("rotation must be (B*n, 4)")
        dev = rotation.device
        q = rotation.detach().float().contiguous()
        B = q.shape[0] // n
        scratch = torch.empty(lib.cg_curve_smooth_scratch_bytes(), dtype=torch.uint8, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        with _lib.on_device(dev):
            _lib.check(lib.cg_curve_smooth_fwd(B, n, _lib.ptr(q), scratch.data_ptr(), loss.data_ptr(),
                                               _lib.stream(dev)), "cg_curve_smooth_fwd")
        ctx.save_for_backward(q)
        ctx.meta = (B, n)
        return loss

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        (q,) = ctx.saved_tensors
        B, n = ctx.meta
        out = torch.empty_like(q)
        g = g.float().contiguous()
        with _lib.on_device(q.device):
            _lib.check(lib.cg_curve_smooth_bwd(B, n, _lib.ptr(q), g.data_ptr(), _lib.ptr(out), _lib.stream(q.device)),
                       "cg_curve_smooth_bwd")
        return out, None


def curve_smoothness(rotation, n_gaussians):
    """Device scalar; `rotation` is the curve model's raw `_rotation` (B*n, 4)."""
    return _CurveSmooth.apply(rotation, int(n_gaussians))


class _EndpointConn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, curve_points, dis_thr):
        lib = _lib.load()
        if not curve_points.is_cuda:
            raise _lib.CurveGSError("endpoint_connectivity needs CUDA tensors; there is no CPU path")
        if curve_points.ndim != 3 or tuple(curve_points.shape[1:]) != (4, 3):
            raise _lib.CurveGSError("curve_points must be (B, 4, 3)")
        dev = curve_points.device
        cp = curve_points.detach().float().contiguous()
        B = cp.shape[0]
        scratch = torch.empty(lib.cg_endpoint_conn_scratch_bytes(B), dtype=torch.uint8, device=dev)
        v = torch.empty((2 * B, 3), dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        with _lib.on_device(dev):
            _lib.check(lib.cg_endpoint_conn_fwd(B, _lib.ptr(cp), float(dis_thr), scratch.data_ptr(), _lib.ptr(v),
                                                loss.data_ptr(), _lib.stream(dev)), "cg_endpoint_conn_fwd")
        ctx.save_for_backward(v, scratch[:64])
        ctx.B = B
        return loss

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        v, sums = ctx.saved_tensors
        B = ctx.B
        out = torch.empty((B, 4, 3), dtype=torch.float32, device=v.device)
        g = g.float().contiguous()
        with _lib.on_device(v.device):
            _lib.check(lib.cg_endpoint_conn_bwd(B, _lib.ptr(v), sums.data_ptr(), g.data_ptr(), _lib.ptr(out),
                                                _lib.stream(v.device)), "cg_endpoint_conn_bwd")
        return out, None


def endpoint_connectivity(curve_points, dis_thr=0.05):
    """Device scalar; `curve_points` is (B,4,3)."""
    return _EndpointConn.apply(curve_points, float(dis_thr))
