"""One training iteration as the reference's train.py runs it (train.py:75-243), on the B200-native ops.

The caller of the hot path: learning-rate update -> render one view -> image loss + curve regularisers ->
backward -> densification statistics / periodic curve-set surgery -> Adam step -> re-sample. Differences
from the reference loop are host-side only:
  * the image loss, curve smoothness and endpoint connectivity are the fused ops (loss.py, regularizers.py);
  * nothing reads a device scalar per iteration (the reference's five `.item()` calls for its progress bar,
    train.py:153-157, and the `visibility_filter.sum() > 0` tests, :114/:119); `stats()` reads them on demand;
  * `merge_curves` (train.py:215) runs in its deterministic, batched form (topology.py): no RANSAC draw.
"""
from __future__ import annotations

import random
from typing import Optional, Sequence

import torch

from . import rasterizer as _rz
from .graph import GraphedStep, StaticCamera
from .loss import edge_ssim_loss
from .parallel import FlatGrad
from .regularizers import curve_smoothness, endpoint_connectivity
from .renderer import render


class OptimizationParams:
    """Defaults of arguments/__init__.py:78-124 (the fields this loop reads)."""
    iterations = 10_000
    position_lr_delay_mult = 0.01
    position_lr_max_steps = 30_000
    lr_curve_points_init = 0.0005
    lr_curve_points_final = 0.000005
    feature_lr = 0.0025
    opacity_lr = 0.025
    scaling_lr = 0.005
    mask_lr = 0.01
    lambda_dssim = 0.1
    opacity_cull = 0.01
    opacity_cull_second = 0.05
    opacity_loss_weight = 0.01
    lambda_mse = 10.0
    lambda_curve_smo = 0.1
    lambda_points_conn = 0.1
    lambda_width = 0.01
    lambda_mask = 0.0005
    mask_threshold = 0.01
    densification_interval = 2000
    opacity_reset_interval = 3000
    densify_from_iter = 500
    densify_until_iter = 7000
    conn_from_iter = 7000
    densify_grad_threshold = 2000
    threshold_line = 0.0015
    threshold_max_line = 0.005
    threshold_angle = 20
    threshold_angle_skip = 30
    distance_threshold = 0.02
    similarity_threshold = 0.97

    def __init__(self, **overrides):
        for k, v in overrides.items():
            if not hasattr(type(self), k):
                raise AttributeError(f"unknown optimization parameter {k}")
            setattr(self, k, v)


class PipelineParams:
    debug = False
    antialiasing = False
    render_geo = True


def opacity_regulariser(opacity: torch.Tensor, visible: torch.Tensor) -> torch.Tensor:
    """mean over the visible Gaussians of log(1 + o^2 / 0.5) (train.py:114-117), 0 when none is visible. A masked
    mean: the same value and gradient as the reference's `opacity[visibility_filter]` without the host-synchronising
    boolean indexing and `.sum() > 0` test."""
    vis = visible.view(-1, 1).to(opacity.dtype)
    return (torch.log(1 + opacity ** 2 / 0.5) * vis).sum() / vis.sum().clamp_min(1.0)


def width_regulariser(width: torch.Tensor, width_thr: float = 0.005) -> torch.Tensor:
    """mean excess over width_thr of the curves at least that wide (train.py:126-131), 0 when there is none."""
    over = (width >= width_thr).to(width.dtype)
    return ((width - width_thr) * over).sum() / over.sum().clamp_min(1.0)


class TrainLoop:
    """`TrainLoop(model, cameras, targets, opt).step()` = one iteration of train.py's loop body.

    cameras: objects with the reference Camera attributes; targets: the per-view edge maps (1,H,W) on the device
    (train.py:99 `original_image[:1]`). `model.training_setup(opt)` is called here.
    """

    def __init__(self, model, cameras: Sequence, targets: Sequence[torch.Tensor], opt: Optional[OptimizationParams] = None,
                 pipe=None, background: Optional[torch.Tensor] = None, cameras_extent: float = 1.0, seed: int = 0,
                 graph: bool = False):
        self.model, self.cameras, self.targets = model, list(cameras), list(targets)
        self.opt = opt or OptimizationParams()
        self.pipe = pipe or PipelineParams()
        dev = model._curve_points.device
        self.bg = background if background is not None else torch.zeros(3, device=dev)
        self.cameras_extent = cameras_extent
        self.iteration = 0
        self.rng = random.Random(seed)
        self._stack: list = []
        model.training_setup(self.opt)
        self.last = {}
        # graph=True: render -> loss -> backward is replayed from a CUDA graph (graph.GraphedStep) and re-captured
        # whenever the curve set or the set of active loss terms changes; all views must share size and FoV
        self.graph = bool(graph)
        self.captures = 0
        self.overflows = 0          # graph mode: how often a periodic check found replays over the captured capacity
        self.verify_every = 64
        self._gs = self._sig = self._fg = None
        if self.graph:
            self._policy = _rz.CapacityBinning()
            self._scam = StaticCamera(self.cameras[0])
            self._gt = torch.empty_like(self.targets[0])

    def _next_view(self) -> int:
        if not self._stack:
            self._stack = list(range(len(self.cameras)))
        return self._stack.pop(self.rng.randint(0, len(self._stack) - 1))

    def loss_terms(self, pkg, gt, iteration):
        """The scalar of train.py:101-146 as device tensors: (total, {name: term})."""
        opt, m = self.opt, self.model
        terms = {"image": edge_ssim_loss(pkg["render_raw"], gt, threshold=0.1, lambda_mse=opt.lambda_mse,
                                         lambda_dssim=opt.lambda_dssim, clamp=True)}
        if iteration >= opt.densify_until_iter:
            terms["mask"] = opt.lambda_mask * torch.sigmoid(m._mask).mean()
        terms["opacity"] = opt.opacity_loss_weight * opacity_regulariser(m.get_opacity, pkg["radii"] > 0)
        if opt.lambda_curve_smo > 0:
            # train.py:119 adds it only `if visibility_filter.sum() > 0`: the same gate, kept on the device
            gate = (pkg["radii"] > 0).any().to(torch.float32)
            terms["curve_smo"] = opt.lambda_curve_smo * gate * curve_smoothness(m._rotation, m.n_gaussians)
        if opt.lambda_width > 0:
            terms["width"] = opt.lambda_width * width_regulariser(m.get_curve_width)
        if opt.lambda_points_conn > 0 and iteration > opt.conn_from_iter:
            terms["curve_conn"] = opt.lambda_points_conn * endpoint_connectivity(m.get_curve_points, 0.05)
        total = terms["image"]
        for k, v in terms.items():
            if k != "image":
                total = total + v
        return total, terms

    def step(self):
        opt, m = self.opt, self.model
        self.iteration += 1
        it = self.iteration
        m.update_learning_rate(it)
        v = self._next_view()
        cam, gt = self.cameras[v], self.targets[v]
        if self.graph:
            loss, terms, pkg = self._graph_forward_backward(cam, gt, it)
        else:
            pkg = render(cam, m, self.pipe, self.bg, use_mask=it >= opt.densify_until_iter, mask_thr=opt.mask_threshold)
            loss, terms = self.loss_terms(pkg, gt, it)
            loss.backward()
        self.last = {"loss": loss.detach(), **{k: t.detach() for k, t in terms.items()}}
        # curve-set edits below replace the parameters: their pending gradients go with the old tensors, exactly as
        # in the reference (the Adam step that follows sees .grad = None for the new tensors and skips them)
        with torch.no_grad():
            radii = pkg["radii"]
            if it < opt.densify_until_iter:
                visible = radii > 0
                m.max_radii2D = torch.where(visible, torch.maximum(m.max_radii2D, radii.to(m.max_radii2D.dtype)), m.max_radii2D)
                m.add_densification_stats(pkg["viewspace_points"], visible)
                if it > opt.densify_from_iter and it % opt.densification_interval == 0:
                    size_threshold = 20 if it > opt.opacity_reset_interval else None
                    m.densify_and_prune(opt.densify_grad_threshold, opt.opacity_cull, self.cameras_extent,
                                        size_threshold, radii)
            if it == opt.densify_until_iter:
                m.prune_curves((m.get_curve_opacity <= opt.opacity_cull_second).reshape(-1))
                m.fix_opacity()
            if it % 1000 == 500 and it > opt.densify_until_iter:
                m.only_prune(opt.opacity_cull, opt.mask_threshold)
                m.mask_trim_split(opt.mask_threshold)
            if it % 1000 == 0 and it > 3000 and it != opt.iterations:
                m.curve_split_curvature(opt.threshold_angle, opt.threshold_angle_skip)
            if (it % 1000 == 0 and it > opt.densify_until_iter) or it == opt.iterations:
                m.fit_curve_to_line(opt.threshold_line, opt.threshold_max_line)
                m.merge_curves(opt.distance_threshold, opt.similarity_threshold)
            if it < opt.iterations:
                m.optimizer.step()
                if not self.graph:      # (the captured step zeroes its flat gradient buffer itself)
                    m.optimizer.zero_grad(set_to_none=True)
        if not self.graph or self._signature(it + 1) != self._sig:
            m.prepare_scaling_rot()     # (a captured step re-samples at its start)
        return loss

    # ---- CUDA-graph mode -------------------------------------------------------------------------------------
    def _trainable(self):
        m = self.model
        return [p for p in (m._curve_points, m._width, m._opacity, m._mask) if p.requires_grad]

    def _signature(self, it):
        opt = self.opt
        m = self.model
        return (getattr(m, "topology_version", 0), tuple(p.requires_grad for p in (m._curve_points, m._width, m._opacity, m._mask)),
                it >= opt.densify_until_iter, opt.lambda_points_conn > 0 and it > opt.conn_from_iter)

    def _graph_forward_backward(self, cam, gt, it):
        m, opt = self.model, self.opt
        sig = self._signature(it)
        if sig != self._sig:
            use_mask = it >= opt.densify_until_iter
            self._gs = None   # the old capture's outputs keep its autograd graph (and that graph's streams) alive

            def body():
                m.prepare_scaling_rot()
                pkg = render(self._scam, m, self.pipe, self.bg, use_mask=use_mask, mask_thr=opt.mask_threshold)
                loss, terms = self.loss_terms(pkg, self._gt, it)
                loss.backward()
                return loss, terms, pkg

            # which tensors autograd reaches with this set of loss terms: only those get a slot in the flat gradient
            # buffer, so Adam skips the others exactly as it does in the eager loop (their .grad stays None)
            params = self._trainable()
            for p in (m._curve_points, m._width, m._opacity, m._mask, m._features_dc, m._features_rest):
                p.grad = None
            self._scam.load(cam)
            self._gt.copy_(gt)
            body()
            reached = [p for p in params if p.grad is not None]
            self._fg = FlatGrad(reached, direct=True)

            def fn():
                self._fg.zero()
                return body()

            loads = [(lambda c=c, t=t: (self._scam.load(c), self._gt.copy_(t))) for c, t in zip(self.cameras, self.targets)]
            self._gs = GraphedStep(fn, policy=self._policy, calibrate=loads).capture()
            self._sig = sig
            self.captures += 1
        self._scam.load(cam)
        self._gt.copy_(gt, non_blocking=True)
        out = self._gs.replay()
        if it % self.verify_every == 0:
            # a view with more tile-instances than the captured capacity renders from a truncated list (the farthest
            # instances are dropped); look at the counters now and then and re-capture with more room if it happened
            torch.cuda.current_stream().synchronize()
            if not self._gs.verify():
                self.overflows += 1
        return out

    def verify(self) -> bool:
        """Graph mode, after a synchronisation: False if recent replays overflowed the captured binning capacity
        (those iterations used truncated instance lists); the step has been re-captured with a larger one."""
        return True if self._gs is None else self._gs.verify()

    def stats(self) -> dict:
        """Host copies of the last iteration's loss terms (one synchronisation, on demand)."""
        return {k: float(v) for k, v in self.last.items()} | {"curves": int(self.model._curve_points.shape[0]),
                                                             "iteration": self.iteration}
