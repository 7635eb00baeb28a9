"""Drop-in for the reference `fused_ssim` package (submodules/fused-ssim/fused_ssim/__init__.py:8-41)."""
from __future__ import annotations

import torch

from . import _lib

allowed_padding = ["same", "valid"]


def _fusedssim(C1, C2, img1, img2, train):
    lib = _lib.load()
    if not img1.is_cuda:
        raise _lib.CurveGSError("fused_ssim needs CUDA tensors; there is no CPU path")
    a = img1.float().contiguous()
    b = img2.float().contiguous()
    B, CH, H, W = a.shape
    m = torch.empty_like(a)
    d1 = torch.empty_like(a) if train else None
    d2 = torch.empty_like(a) if train else None
    d3 = torch.empty_like(a) if train else None
    with _lib.on_device(a.device):
        _lib.check(lib.cg_ssim_fwd(B, CH, H, W, float(C1), float(C2), _lib.ptr(a), _lib.ptr(b), _lib.ptr(m),
                                   _lib.ptr(d1), _lib.ptr(d2), _lib.ptr(d3),
                                   _lib.stream(a.device)), "cg_ssim_fwd")
    e = torch.empty(0)
    return m, (d1 if train else e), (d2 if train else e), (d3 if train else e)


def _fusedssim_backward(C1, C2, img1, img2, dL_dmap, d1, d2, d3):
    lib = _lib.load()
    a = img1.float().contiguous()
    b = img2.float().contiguous()
    g = dL_dmap.float().contiguous()
    B, CH, H, W = a.shape
    out = torch.empty_like(a)
    with _lib.on_device(a.device):
        _lib.check(lib.cg_ssim_bwd(B, CH, H, W, float(C1), float(C2), _lib.ptr(a), _lib.ptr(b), _lib.ptr(g),
                                   _lib.ptr(d1), _lib.ptr(d2), _lib.ptr(d3), _lib.ptr(out),
                                   _lib.stream(a.device)), "cg_ssim_bwd")
    return out


class FusedSSIMMap(torch.autograd.Function):
    @staticmethod
    def forward(ctx, C1, C2, img1, img2, padding="same", train=True):
        ssim_map, dm_dmu1, dm_dsigma1_sq, dm_dsigma12 = _fusedssim(C1, C2, img1, img2, train)
        if padding == "valid":
            ssim_map = ssim_map[:, :, 5:-5, 5:-5]
        ctx.save_for_backward(img1.detach(), img2, dm_dmu1, dm_dsigma1_sq, dm_dsigma12)
        ctx.C1 = C1
        ctx.C2 = C2
        ctx.padding = padding
        return ssim_map

    @staticmethod
    def backward(ctx, opt_grad):
        img1, img2, dm_dmu1, dm_dsigma1_sq, dm_dsigma12 = ctx.saved_tensors
        C1, C2, padding = ctx.C1, ctx.C2, ctx.padding
        dL_dmap = opt_grad
        if padding == "valid":
            dL_dmap = torch.zeros_like(img1)
            dL_dmap[:, :, 5:-5, 5:-5] = opt_grad
        grad = _fusedssim_backward(C1, C2, img1, img2, dL_dmap, dm_dmu1, dm_dsigma1_sq, dm_dsigma12)
        return None, None, grad, None, None, None


def fused_ssim(img1, img2, padding="same", train=True):
    C1 = 0.01 ** 2
    C2 = 0.03 ** 2
    assert padding in allowed_padding
    return FusedSSIMMap.apply(C1, C2, img1, img2, padding, train).mean()
