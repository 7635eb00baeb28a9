"""Fused per-view activations (one CUDA kernel forward, one backward) used by render().

Numerically this is the chain the reference spells in ~60 ATen ops per view:
get_rotation / get_opacity / mask straight-through (gaussian_renderer/__init__.py:57-76),
get_main_axis + view-space rotation + all_map assembly (:98-104, gaussian_curve_model.py:99-105).
"""
from __future__ import annotations

import torch

from . import _lib


class _CurveActivate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, rotation, scaling, opacity_logit, mask_logit, n, mask_thr, campos, viewmatrix,
                direct_params=None):
        lib = _lib.load()
        if not xyz.is_cuda:
            raise _lib.CurveGSError("activation needs CUDA tensors; there is no CPU path")
        dev = xyz.device
        P = xyz.shape[0]
        B = P // n
        f = lambda t: t.detach().float().contiguous()
        xyz_, rot_, scal_ = f(xyz), f(rotation), f(scaling)
        ol_ = f(opacity_logit).view(-1)
        ml_ = f(mask_logit).view(-1) if mask_logit is not None else None
        cam_, vm_ = f(campos), f(viewmatrix)
        rot_n = torch.empty((P, 4), dtype=torch.float32, device=dev)
        opacity = torch.empty((P, 1), dtype=torch.float32, device=dev)
        scales = torch.empty((P, 3), dtype=torch.float32, device=dev)
        all_map = torch.empty((P, 4), dtype=torch.float32, device=dev)
        with _lib.on_device(dev):
            _lib.check(lib.cg_activate_fwd(B, n, _lib.ptr(xyz_), _lib.ptr(rot_), _lib.ptr(scal_), _lib.ptr(ol_),
                                           _lib.ptr(ml_), float(mask_thr), _lib.ptr(cam_), _lib.ptr(vm_),
                                           _lib.ptr(rot_n), _lib.ptr(opacity), _lib.ptr(scales), _lib.ptr(all_map),
                                           _lib.stream(dev)), "cg_activate_fwd")
        ctx.save_for_backward(xyz_, rot_, scal_, ol_, ml_ if ml_ is not None else torch.empty(0, device=dev), cam_, vm_)
        ctx.meta = (B, n, float(mask_thr), tuple(opacity_logit.shape),
                    tuple(mask_logit.shape) if mask_logit is not None else None)
        ctx.set_materialize_grads(False)
        ctx.direct = direct_params     # (opacity, mask) parameter objects when they arrived detached: curve_activate
        return rot_n, opacity, scales, all_map

    @staticmethod
    def backward(ctx, g_rot_n, g_opacity, g_scales, g_all_map):
        lib = _lib.load()
        xyz_, rot_, scal_, ol_, ml_, cam_, vm_ = ctx.saved_tensors
        B, n, thr, ol_shape, ml_shape = ctx.meta
        dev = xyz_.device
        P = B * n
        c = lambda g: None if g is None else g.float().contiguous()
        g_rot_n, g_opacity, g_scales, g_all_map = c(g_rot_n), c(g_opacity), c(g_scales), c(g_all_map)
        g_rot = torch.empty((P, 4), dtype=torch.float32, device=dev)
        g_scaling = torch.empty((P, 3), dtype=torch.float32, device=dev)
        from .sampling import direct_grad
        direct = ctx.direct is not None and all(p is None or direct_grad(p) is not None for p in ctx.direct)
        if ctx.direct is not None and not direct:
            raise _lib.CurveGSError("a parameter's FlatGrad(direct=True) gradient buffer was replaced between forward and "
                                    "backward (call FlatGrad.bind() after an optimizer step that set .grad to None)")
        if direct:
            g_ol = direct_grad(ctx.direct[0])
            g_ml = direct_grad(ctx.direct[1]) if ml_shape is not None else None
        else:
            g_ol = torch.empty((B,), dtype=torch.float32, device=dev)
            g_ml = torch.empty((P,), dtype=torch.float32, device=dev) if ml_shape is not None else None
        with _lib.on_device(dev):
            _lib.check(lib.cg_activate_bwd(B, n, _lib.ptr(xyz_), _lib.ptr(rot_), _lib.ptr(scal_), _lib.ptr(ol_),
                                           _lib.ptr(ml_) if ml_shape is not None else None, thr, _lib.ptr(cam_),
                                           _lib.ptr(vm_), _lib.ptr(g_rot_n), _lib.ptr(g_opacity), _lib.ptr(g_scales),
                                           _lib.ptr(g_all_map), _lib.ptr(g_rot), _lib.ptr(g_scaling), _lib.ptr(g_ol),
                                           _lib.ptr(g_ml), 1 if direct else 0, _lib.stream(dev)),
                       "cg_activate_bwd")
        if direct:
            return (None, g_rot, g_scaling, None, None, None, None, None, None, None)
        return (None, g_rot, g_scaling, g_ol.view(ol_shape), g_ml.view(ml_shape) if g_ml is not None else None,
                None, None, None, None, None)


def curve_activate(xyz, rotation, scaling, opacity_logit, mask_logit, n, campos, viewmatrix, use_mask=False,
                   mask_thr=0.01):
    """-> rotations (P,4) unit, opacity (P,1), scales (P,3), all_map (P,4)."""
    from .sampling import direct_grad, direct_target
    ml = mask_logit if use_mask else None
    if torch.is_grad_enabled() and direct_target(opacity_logit) and direct_grad(opacity_logit) is not None and \
            (ml is None or (direct_target(ml) and direct_grad(ml) is not None)) and \
            (rotation.requires_grad or scaling.requires_grad):
        # direct mode (parallel.FlatGrad(direct=True)): the parameters go in detached and the backward kernel adds
        # their gradients into .grad itself - no edge to their AccumulateGrad nodes (see sampling.sample_curves)
        return _CurveActivate.apply(xyz, rotation, scaling, opacity_logit.detach(), ml.detach() if ml is not None else None,
                                    int(n), float(mask_thr), campos, viewmatrix, (opacity_logit, ml))
    return _CurveActivate.apply(xyz, rotation, scaling, opacity_logit, ml, int(n), float(mask_thr), campos, viewmatrix)
