"""Fused image loss of the training step and the direction-channel rotation of render().

`edge_ssim_loss(image, gt)` equals, for single-channel images, what train.py:101-107 spells as

    Ll1 = edge_aware_loss(image, gt)                      (utils/loss_utils.py:94-115)
    ssim_value = fused_ssim(image[None], gt[None])        (fused_ssim/__init__.py:34-41)
    loss = lambda_mse * ((1 - lambda_dssim) * Ll1 + lambda_dssim * (1 - ssim_value))

but runs as ONE forward and ONE backward CUDA kernel (cg_edge_ssim_loss_fwd/bwd) and never
synchronises with the host: the scalar stays on the device. `edge_aware_loss` below is the
reference's own torch expression, kept for callers that want the two terms separately.
"""
from __future__ import annotations

import torch

from . import _lib


def edge_aware_loss(image, gt_image, threshold=0.1):
    """The reference's class-balanced weighted MSE, verbatim semantics (utils/loss_utils.py:94-115)."""
    edge_map = gt_image.mean(dim=0, keepdim=True)
    num_positive = (torch.sum(edge_map > threshold)).float()
    num_negative = (torch.sum(edge_map <= threshold)).float()
    mask = torch.where(edge_map > threshold, 5. * (num_negative + 1) / (num_positive + num_negative),
                       1.0 * (num_positive + 1) / (num_positive + num_negative))
    return (((image - gt_image) ** 2) * mask).mean()


class _EdgeSSIMLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, gt, threshold, lambda_mse, lambda_dssim, clamp):
        lib = _lib.load()
        if not image.is_cuda:
            raise _lib.CurveGSError("edge_ssim_loss needs CUDA tensors; there is no CPU path")
        if image.numel() != image.shape[-1] * image.shape[-2] or gt.numel() != image.numel():
            raise _lib.CurveGSError("edge_ssim_loss is defined for single-channel (1,H,W) images of equal size")
        dev = image.device
        H, W = int(image.shape[-2]), int(image.shape[-1])
        a = image.detach().float().contiguous()
        b = gt.detach().float().contiguous()
        stats = torch.empty(lib.cg_edge_ssim_loss_stats_bytes() // 8, dtype=torch.float64, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        train = image.requires_grad
        d1 = torch.empty_like(a) if train else None
        d2 = torch.empty_like(a) if train else None
        d3 = torch.empty_like(a) if train else None
        with _lib.on_device(dev):
            _lib.check(lib.cg_edge_ssim_loss_fwd(H, W, a.data_ptr(), b.data_ptr(), float(threshold), float(lambda_mse),
                                                 float(lambda_dssim), 0.01 ** 2, 0.03 ** 2, int(bool(clamp)), stats.data_ptr(),
                                                 loss.data_ptr(), _lib.ptr(d1), _lib.ptr(d2), _lib.ptr(d3),
                                                 _lib.stream(dev)), "cg_edge_ssim_loss_fwd")
        if train:
            ctx.save_for_backward(a, b, stats, d1, d2, d3)
        ctx.meta = (H, W, float(threshold), float(lambda_mse), float(lambda_dssim), int(bool(clamp)), tuple(image.shape))
        return loss

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        a, b, stats, d1, d2, d3 = ctx.saved_tensors
        H, W, thr, lm, ld, clamp, shape = ctx.meta
        out = torch.empty_like(a)
        g = g.float().contiguous()
        with _lib.on_device(a.device):
            _lib.check(lib.cg_edge_ssim_loss_bwd(H, W, a.data_ptr(), b.data_ptr(), thr, lm, ld, clamp, stats.data_ptr(),
                                                 g.data_ptr(), d1.data_ptr(), d2.data_ptr(), d3.data_ptr(),
                                                 out.data_ptr(), _lib.stream(a.device)),
                       "cg_edge_ssim_loss_bwd")
        return out.view(shape), None, None, None, None, None


def edge_ssim_loss(image, gt_image, threshold=0.1, lambda_mse=10.0, lambda_dssim=0.1, clamp=False):
    """Device scalar lambda_mse*((1-lambda_dssim)*edge_aware_loss + lambda_dssim*(1-ssim)); defaults are the
    reference's OptimizationParams (arguments/__init__.py:94-110). With clamp=True `image` is the raw render
    (`render(...)["render_raw"]`) and render()'s clamp(0,1) is applied inside the kernels, forward and adjoint."""
    return _EdgeSSIMLoss.apply(image, gt_image, threshold, lambda_mse, lambda_dssim, clamp)


class _RotateChannels(torch.autograd.Function):
    """(3,H,W) planar image, out[c] = sum_k in[k] * M[c][k] (M a 3x3 view of a device matrix)."""

    @staticmethod
    def forward(ctx, planes, m):
        lib = _lib.load()
        if not planes.is_cuda:
            raise _lib.CurveGSError("rotate_channels needs CUDA tensors; there is no CPU path")
        x = planes.float().contiguous()
        mm = m.detach().float()
        if mm.stride(-1) != 1:
            mm = mm.contiguous()
        n = x.shape[-1] * x.shape[-2]
        out = torch.empty_like(x)
        with _lib.on_device(x.device):
            _lib.check(lib.cg_rotate_channels(n, x.data_ptr(), mm.data_ptr(), int(mm.stride(0)), 0, out.data_ptr(),
                                              _lib.stream(x.device)), "cg_rotate_channels")
        ctx.save_for_backward(mm)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        (mm,) = ctx.saved_tensors
        g = g.float().contiguous()
        n = g.shape[-1] * g.shape[-2]
        out = torch.empty_like(g)
        with _lib.on_device(g.device):
            _lib.check(lib.cg_rotate_channels(n, g.data_ptr(), mm.data_ptr(), int(mm.stride(0)), 1, out.data_ptr(),
                                              _lib.stream(g.device)), "cg_rotate_channels")
        return out, None


def rotate_channels(planes, m3x3):
    """Equals (planes.permute(1,2,0) @ m3x3.T).permute(2,0,1) for a (3,H,W) tensor."""
    return _RotateChannels.apply(planes, m3x3)
