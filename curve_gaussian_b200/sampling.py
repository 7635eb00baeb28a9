"""Curve control points -> sampled Gaussians as ONE fused forward/backward op.

Replaces the ~50 ATen launches (+ autograd graph) of
GaussianCurveModel.prepare_scaling_rot (scene/gaussian_curve_model.py:180-198)
and rot_to_quat_batch (utils/general_utils.py:33-86) with cg_sample_fwd /
cg_sample_bwd from libcurvegs.so. Outputs and gradients follow the reference,
including the two whole-tensor norms that couple all Gaussians.
"""
from __future__ import annotations

import torch

from . import _lib


def sample_t(n: int, device) -> torch.Tensor:
    """The reference's sample parameters (gaussian_curve_model.py:58-60), shape (n,)."""
    return torch.linspace(0.5 / n, 1 - 0.5 / n, n, device=device)


def direct_target(p) -> bool:
    """Is `p` a leaf parameter bound to a FlatGrad(direct=True) buffer (parallel.py)?"""
    return bool(getattr(p, "_cg_direct_grad", False)) and p.is_leaf and p.requires_grad


def direct_grad(p):
    """The gradient buffer a backward kernel may add into, or None (then the gradient is returned to autograd)."""
    g = p.grad
    if g is None or g.dtype != torch.float32 or not g.is_contiguous() or g.shape != p.shape or g.device != p.device:
        return None
    return g


class _CurveSample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, curve_points, width, is_bezier, t, anchor=None, direct_params=None):
        lib = _lib.load()
        if not curve_points.is_cuda:
            raise _lib.CurveGSError("curve_points must be a CUDA tensor; there is no CPU path")
        dev = curve_points.device
        B = curve_points.shape[0]
        n = t.numel()
        cp = curve_points.detach().float().contiguous()
        w = width.detach().float().contiguous().view(-1)
        isb = is_bezier.to(device=dev, dtype=torch.uint8).contiguous() if is_bezier is not None else None
        tt = t.detach().float().contiguous().view(-1)
        P = B * n
        xyz = torch.empty((P, 3), dtype=torch.float32, device=dev)
        rot = torch.empty((P, 4), dtype=torch.float32, device=dev)
        scaling = torch.empty((P, 3), dtype=torch.float32, device=dev)
        norms = torch.empty(2, dtype=torch.float32, device=dev)
        scratch = torch.empty(128, dtype=torch.uint8, device=dev)   # the forward only needs the 64-byte sums block
        half_step = 0.5 / n
        with _lib.on_device(dev):
            _lib.check(lib.cg_sample_fwd(B, n, _lib.ptr(cp), _lib.ptr(w), _lib.ptr(isb), _lib.ptr(tt), half_step,
                                         _lib.ptr(xyz), _lib.ptr(rot), _lib.ptr(scaling), norms.data_ptr(),
                                         scratch.data_ptr(), _lib.stream(dev)),
                       "cg_sample_fwd")
        ctx.save_for_backward(cp, w, isb if isb is not None else torch.empty(0, device=dev), tt, norms)
        ctx.shape = (B, n, half_step, tuple(width.shape))
        # parameters whose .grad is a view of a FlatGrad(direct=True) buffer get their gradient ADDED there by the
        # backward kernel itself (no AccumulateGrad node, no ATen add / fill kernels): see parallel.FlatGrad
        # (they arrive DETACHED, next to the parameter objects themselves and a fresh `anchor` leaf that keeps the
        # outputs differentiable: see sample_curves)
        ctx.direct = direct_params
        ctx.set_materialize_grads(False)
        return xyz, rot, scaling

    @staticmethod
    def backward(ctx, g_xyz, g_rot, g_scaling):
        lib = _lib.load()
        cp, w, isb, tt, norms = ctx.saved_tensors
        B, n, half_step, wshape = ctx.shape
        dev = cp.device
        direct = ctx.direct is not None and all(direct_grad(p) is not None for p in ctx.direct)
        if ctx.direct is not None and not direct:
            raise _lib.CurveGSError("a parameter's FlatGrad(direct=True) gradient buffer was replaced between forward and "
                                    "backward (call FlatGrad.bind() after an optimizer step that set .grad to None)")
        if direct:
            g_cp, g_w = direct_grad(ctx.direct[0]), direct_grad(ctx.direct[1])
        else:
            g_cp = torch.empty((B, 4, 3), dtype=torch.float32, device=dev)
            g_w = torch.empty((B,), dtype=torch.float32, device=dev)
        scratch = torch.empty(max(lib.cg_sample_scratch_bytes(B, n), 8), dtype=torch.uint8, device=dev)
        c = lambda g: None if g is None else g.float().contiguous()
        g_xyz, g_rot, g_scaling = c(g_xyz), c(g_rot), c(g_scaling)
        with _lib.on_device(dev):
            _lib.check(lib.cg_sample_bwd(B, n, _lib.ptr(cp), _lib.ptr(w), _lib.ptr(isb), _lib.ptr(tt), half_step,
                                         norms.data_ptr(), _lib.ptr(g_xyz), _lib.ptr(g_rot), _lib.ptr(g_scaling),
                                         _lib.ptr(g_cp), _lib.ptr(g_w), scratch.data_ptr(), 1 if direct else 0,
                                         _lib.stream(dev)), "cg_sample_bwd")
        if direct:
            return None, None, None, None, None, None
        return g_cp, g_w.view(wshape), None, None, None, None


def sample_curves(curve_points, width, is_bezier, t):
    """(B,4,3), (B,1), (B,) bool, (n,) -> xyz (B*n,3), rotation (B*n,4), scaling (B*n,3); g = b*n + m."""
    if curve_points.shape[0] == 0:
        z = lambda k: curve_points.new_zeros((0, k))
        return z(3), z(4), z(3)
    if torch.is_grad_enabled() and direct_target(curve_points) and direct_target(width) \
            and direct_grad(curve_points) is not None and direct_grad(width) is not None:
        # Direct mode: the op gets the parameters detached, so the autograd graph has no edge to their AccumulateGrad
        # nodes at all. Those nodes remember the stream they were created on, and every backward that reaches one
        # makes the calling stream wait for that stream - a cross-stream dependency that invalidates a CUDA-graph
        # capture taken on any other stream (cudaErrorStreamCaptureIsolation / ...Implicit). A fresh scalar leaf made
        # on the current stream keeps the outputs differentiable; its gradient is None.
        anchor = torch.empty((), dtype=torch.float32, device=curve_points.device, requires_grad=True)
        return _CurveSample.apply(curve_points.detach(), width.detach(), is_bezier, t, anchor, (curve_points, width))
    return _CurveSample.apply(curve_points, width, is_bezier, t)
