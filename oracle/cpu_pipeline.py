"""TEST INFRASTRUCTURE ONLY — the whole hot path on host cores: torch (CPU) restatement of
the sampling / activations / losses around the C oracle rasterizer. Used as the checker in
tests/ and smoke(), and as the timed CPU baseline in bench.py (cpu_baseline, --impl reference).
"""
from __future__ import annotations

import math
import time

import numpy as np
import torch

from . import cpu as O
from . import torch_ref


def edge_aware_loss(image, gt_image, threshold=0.1):
    """utils/loss_utils.py:94-115 (caller-side loss, restated)."""
    edge_map = gt_image.mean(dim=0, keepdim=True)
    num_positive = (torch.sum(edge_map > threshold)).float()
    num_negative = (torch.sum(edge_map <= threshold)).float()
    mask = torch.zeros_like(edge_map)
    mask[edge_map > threshold] = 5. * (num_negative + 1) / (num_positive + num_negative)
    mask[edge_map <= threshold] = 1.0 * (num_positive + 1) / (num_positive + num_negative)
    loss = (image - gt_image) ** 2
    return (loss * mask).mean()


class _OracleRaster(torch.autograd.Function):
    """CPU stand-in for _RasterizeGaussians built on the C oracle; `tile_rows` restricts the
    blend passes to the first rows of tiles (bounded CPU samples)."""

    @staticmethod
    def forward(ctx, means3D, opac, scales, rots, colors, all_map, cam, W, H, bg, tile_rows, timing):
        t0 = time.perf_counter()
        tanx, tany = math.tan(cam.FoVx * 0.5), math.tan(cam.FoVy * 0.5)
        vm = cam.world_view_transform.numpy()
        pm = cam.full_proj_transform.numpy()
        npy = lambda t: t.detach().contiguous().numpy()
        g = O.preprocess(npy(means3D), npy(scales), npy(rots), npy(opac).reshape(-1), vm, pm, W, H, tanx, tany)
        keys, vals, ranges = O.bin_tiles(g)
        t1 = time.perf_counter()
        R_total = len(keys)
        gx = (W + 15) // 16
        if tile_rows is not None:
            ranges = ranges.copy()
            ranges[tile_rows * gx:] = 0
        R_used = int((ranges[:, 1].astype(np.int64) - ranges[:, 0].astype(np.int64)).sum())
        fwd = O.blend_fwd(g, ranges, vals, npy(colors).reshape(-1), npy(all_map), np.asarray(bg, np.float32), True)
        t2 = time.perf_counter()
        ctx.stuff = (g, ranges, vals, fwd, npy(colors).reshape(-1), npy(all_map), np.asarray(bg, np.float32), vm, pm,
                     tanx, tany, npy(means3D), npy(scales), npy(rots), npy(opac).reshape(-1), timing)
        if timing is not None:
            timing.update(geom_bin_s=t1 - t0, blend_fwd_s=t2 - t1, R_total=R_total, R_used=R_used)
        return torch.from_numpy(fwd.color), torch.from_numpy(fwd.all_map)

    @staticmethod
    def backward(ctx, g_color, g_map):
        (g, ranges, vals, fwd, colors, all_map, bg, vm, pm, tanx, tany, means, scales, rots, opac, timing) = ctx.stuff
        t0 = time.perf_counter()
        acc = O.blend_bwd(g, ranges, vals, colors, all_map, bg, fwd, g_color.contiguous().numpy(), None,
                          None if g_map is None else g_map.contiguous().numpy(), True)
        t1 = time.perf_counter()
        pb = O.preprocess_bwd(g, acc, means, scales, rots, opac, vm, pm, tanx, tany)
        t2 = time.perf_counter()
        if timing is not None:
            timing.update(blend_bwd_s=t1 - t0, preprocess_bwd_s=t2 - t1)
        f = torch.from_numpy
        return (f(pb.d_means3D), f(pb.d_opacity).view(-1, 1), f(pb.d_scales), f(pb.d_rots), None,
                f(acc.d_all_map), None, None, None, None, None, None)


def cpu_train_step(curve_points, width, opacity_logit, mask_logit, is_bezier, n, cam, gt, tile_rows=None,
                   with_ssim=True):
    """One view of the hot path on the CPU: sampling -> activations -> rasterize -> loss -> adjoint back to the
    curve parameters. Returns (loss, grads dict, timing dict)."""
    timing = {}
    t0 = time.perf_counter()
    cp = curve_points.detach().clone().requires_grad_(True)
    w = width.detach().clone().requires_grad_(True)
    ol = opacity_logit.detach().clone().requires_grad_(True)
    xyz, rot, scal = torch_ref.sample_curves(cp, w, is_bezier, n)
    m3, op, scl, rot_n, col, amap = torch_ref.raster_inputs(xyz, rot, scal, ol, n, mask_logit, cam.camera_center,
                                                            cam.world_view_transform)
    t1 = time.perf_counter()
    W, H = cam.image_width, cam.image_height
    color, omap = _OracleRaster.apply(m3, op, scl, rot_n, col, amap, cam, W, H, [0.0, 0.0, 0.0], tile_rows, timing)
    image = color.clamp(0, 1)
    t2 = time.perf_counter()
    Ll1 = edge_aware_loss(image, gt)
    if with_ssim:
        m, d1, d2, d3 = O.ssim_fwd(image[None].detach().numpy(), gt[None].numpy())
        ssim_val = float(m.mean())
        loss = 10.0 * (0.9 * Ll1 + 0.1 * (1.0 - ssim_val))
        g_ssim = O.ssim_bwd(image[None].detach().numpy(), gt[None].numpy(),
                            np.full(m.shape, 1.0 / m.size, np.float32), d1, d2, d3)[0]
        # d loss / d image = 10*0.9*dLl1 - 10*0.1*dssim
        g_l1 = torch.autograd.grad(10.0 * 0.9 * Ll1, image, retain_graph=True)[0]
        g_img = g_l1 - 1.0 * torch.from_numpy(g_ssim)
        t3 = time.perf_counter()
        image.backward(g_img)
    else:
        loss = 10.0 * Ll1
        t3 = time.perf_counter()
        loss.backward()
    t4 = time.perf_counter()
    timing.update(sample_s=t1 - t0, raster_fwd_s=t2 - t1, loss_s=t3 - t2, backward_s=t4 - t3, total_s=t4 - t0)
    return float(loss), dict(curve_points=cp.grad, width=w.grad, opacity=ol.grad), timing


def edge_ssim_loss_ref(image, gt, threshold=0.1, lambda_mse=10.0, lambda_dssim=0.1):
    """CPU restatement of train.py:101-107 for a (1,H,W) image: returns (loss, dL/dimage).
    edge_aware_loss is the torch expression above; SSIM value and gradient come from the C oracle
    (fused-ssim/ssim.cu:187-366 restated in oracle/ssim_oracle.c)."""
    img = image.detach().clone().float().requires_grad_(True)
    Ll1 = edge_aware_loss(img, gt, threshold)
    m, d1, d2, d3 = O.ssim_fwd(img[None].detach().numpy(), gt[None].numpy())
    loss = lambda_mse * ((1.0 - lambda_dssim) * float(Ll1) + lambda_dssim * (1.0 - float(m.mean())))
    g_l1 = torch.autograd.grad(lambda_mse * (1.0 - lambda_dssim) * Ll1, img)[0]
    g_ssim = O.ssim_bwd(img[None].detach().numpy(), gt[None].numpy(), np.full(m.shape, 1.0 / m.size, np.float32),
                        d1, d2, d3)[0]
    return loss, g_l1 - lambda_mse * lambda_dssim * torch.from_numpy(g_ssim)
