"""Drop-in for the reference `diff_cur_rasterization` Python module.

Same names, argument order, return values and error behaviour as
submodules/diff-cur-rasterization/diff_cur_rasterization/__init__.py
(GaussianRasterizationSettings :153-167, GaussianRasterizer :169-222,
_RasterizeGaussians :46-151), but every kernel is libcurvegs.so (sm_100a)
called through the C ABI; torch only owns memory and the stream.
"""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple

import torch
import torch.nn as nn

from . import _lib


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool
    antialiasing: bool
    render_geo: bool


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _settings_struct(rs: GaussianRasterizationSettings, keep: list) -> _lib.RasterSettings:
    bg = _f32c(rs.bg)
    vm = _f32c(rs.viewmatrix)
    pm = _f32c(rs.projmatrix)
    keep.extend([bg, vm, pm])
    for name, t in (("bg", bg), ("viewmatrix", vm), ("projmatrix", pm)):
        if not t.is_cuda:
            raise _lib.CurveGSError(f"raster_settings.{name} must be a CUDA tensor")
    s = _lib.RasterSettings()
    s.image_height = int(rs.image_height)
    s.image_width = int(rs.image_width)
    s.tanfovx = float(rs.tanfovx)
    s.tanfovy = float(rs.tanfovy)
    s.scale_modifier = float(rs.scale_modifier)
    s.render_geo = int(bool(rs.render_geo))
    s.debug = int(bool(rs.debug))
    s.antialiasing = int(bool(rs.antialiasing))
    s.bg = bg.data_ptr()
    s.viewmatrix = vm.data_ptr()
    s.projmatrix = pm.data_ptr()
    return s


def _bytes(n: int, device) -> torch.Tensor:
    # torch's caching allocator returns 512-byte aligned blocks
    return torch.empty(max(int(n), 1), dtype=torch.uint8, device=device)


def _alloc_instances(R: int) -> int:
    """Instance count to SIZE the R-dependent buffers for: R rounded up to the next of {1, 1.25, 1.5, 1.75} x 2^k.
    R changes from view to view; with exact sizes every view that needs more than any before it sends the caching
    allocator to cudaMalloc (a device synchronisation, milliseconds), for as long as new maxima keep arriving. With a
    handful of distinct sizes the blocks of the first few views serve all later ones."""
    R = max(int(R), 1)
    k = max(R.bit_length() - 1, 2)
    for q in (4, 5, 6, 7, 8):
        if R <= (q << (k - 2)):
            return q << (k - 2)
    return R


def _capturing() -> bool:
    return torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()


def _pinned_i32(n: int) -> torch.Tensor:
    t = torch.zeros(n, dtype=torch.int32)
    return t.pin_memory() if torch.cuda.is_available() else t


class CapacityOverflow(_lib.CurveGSError):
    """A sync-free forward met more tile-instances than its binning capacity: that step's outputs are not the
    reference's. The policy has already grown the capacity; repeat the step."""


class CapacityBinning:
    """Sync-free binning policy.

    The reference reads the tile-instance count R back to the host in every forward to size its binning
    buffers (rasterizer_impl.cu:283-291); that round trip is what keeps a small scene host-bound and what
    makes the step impossible to capture in a CUDA graph. With this policy active (`capacity_binning()`),
    the first forward of each (P, W, H) shape runs the exact path to learn R; later ones call
    `cg_raster_fwd_capacity` with buffers sized for `headroom * max R seen`, never touch the host, and leave
    {R, overflow} in a pinned host slot behind the step. `poll()` (non-blocking, call it once per iteration)
    or `check()` (after the caller synchronised anyway) raise `CapacityOverflow` for a step that overflowed
    and keep the capacity tracking the largest R observed.
    """

    def __init__(self, headroom: float = 1.3, granule: int = 1 << 16):
        self.headroom = float(headroom)
        self.granule = int(granule)
        self.caps: dict = {}          # (P, W, H) -> capacity in tile-instances
        self.max_seen: dict = {}      # (P, W, H) -> largest R observed
        self._pending: list = []      # (key, pinned int32[2], event, capacity used)
        self._free: list = []
        self._static: dict = {}       # key -> pinned int32[3] refreshed by graph replays
        self._sticky: dict = {}       # key -> device int32[2]: running max of {R, overflow} over the graph replays
                                      # since the last check() (a later, smaller view must not hide an overflow)
        self.overflows = 0

    def _round(self, R: int) -> int:
        g = self.granule
        return max(g, (int(R * self.headroom) + g - 1) // g * g)

    def learn(self, key, R: int, device=None) -> None:
        if key not in self._static and not _capturing():
            # (pinned allocations are not allowed while a capture is open, so the slot exists beforehand)
            self._static[key] = _pinned_i32(3)
        if device is not None and key not in self._sticky and not _capturing():
            self._sticky[key] = torch.zeros(2, dtype=torch.int32, device=device)
        if R > self.max_seen.get(key, -1):
            self.max_seen[key] = int(R)
            if self._round(R) > self.caps.get(key, 0):
                self.caps[key] = self._round(R)

    def capacity(self, key):
        return self.caps.get(key)

    def observe(self, key, counter: torch.Tensor, cap: int) -> None:
        """Queue the device counter {R, overflow} of a capacity forward for a later poll()/check()."""
        if _capturing():
            # inside a CUDA graph: a fixed pinned slot, rewritten by every replay, read by check()
            slot = self._static[key]
            slot[2] = cap
            sticky = self._sticky.get(key)
            if sticky is None:
                raise _lib.CurveGSError(f"no overflow accumulator for (P, W, H) = {key}: run the step once outside "
                                        "the CUDA graph capture first")
            torch.maximum(sticky, counter, out=sticky)    # part of the graph: every replay folds its {R, overflow} in
            slot[:2].copy_(sticky, non_blocking=True)
            return
        if len(self._pending) >= 64:
            self.poll()           # a caller that never polls: keep the queue bounded (may raise for an earlier step)
        slot = self._free.pop() if self._free else _pinned_i32(2)
        slot.copy_(counter, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._pending.append((key, slot, ev, cap))

    def _account(self, key, R: int, over: int, cap: int) -> bool:
        self.learn(key, R)
        if over:
            self.overflows += 1
        return bool(over)

    def poll(self, block: bool = False) -> None:
        """Account for every finished capacity forward; raise CapacityOverflow if one of them overflowed."""
        bad = None
        keep = []
        for key, slot, ev, cap in self._pending:
            if block:
                ev.synchronize()
            elif not ev.query():
                keep.append((key, slot, ev, cap))
                continue
            R, over = int(slot[0]), int(slot[1])
            self._free.append(slot)
            if self._account(key, R, over, cap):
                bad = (key, R, cap)
        self._pending = keep
        if bad is not None:
            raise CapacityOverflow(f"{bad[1]} tile-instances exceeded the binning capacity {bad[2]} for "
                                   f"(P, W, H) = {bad[0]}; the capacity was raised, repeat the step")

    def check(self) -> None:
        """After the caller synchronised: poll(block=True) plus the slots written by CUDA-graph replays."""
        self.poll(block=True)
        for key, slot in self._static.items():
            R, over, cap = int(slot[0]), int(slot[1]), int(slot[2])
            slot[1] = 0   # reported once
            if key in self._sticky:
                self._sticky[key].zero_()   # stream-ordered behind the replays the caller synchronised on
            if self._account(key, R, over, cap):
                raise CapacityOverflow(f"{R} tile-instances exceeded the captured binning capacity {cap} for "
                                       f"(P, W, H) = {key}; re-capture the step (the capacity was raised)")


_policy: CapacityBinning | None = None


class capacity_binning:
    """Context manager / switch: `with capacity_binning() as pol: ...` or `pol = capacity_binning().enable()`."""

    def __init__(self, policy: CapacityBinning | None = None, **kw):
        self.policy = policy or CapacityBinning(**kw)
        self._prev = None

    def enable(self) -> CapacityBinning:
        global _policy
        self._prev, _policy = _policy, self.policy
        return self.policy

    def disable(self) -> None:
        global _policy
        _policy = self._prev

    def __enter__(self) -> CapacityBinning:
        return self.enable()

    def __exit__(self, *exc):
        self.disable()
        return False


def rasterize_forward_raw(rs, means3D, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, all_maps,
                          capacity=None):
    """The `_C.rasterize_gaussians` equivalent (rasterize_points.cu:35-130).

    Returns (num_rendered, color, radii, geomBuffer, binningBuffer, imgBuffer, invdepths, out_all_map).
    With `capacity` (an int, or the active CapacityBinning policy's choice) the forward does not read R
    back: `num_rendered` is then the CAPACITY the buffers were sized for (pass it on to the backward) and the
    true count is in `rasterize_forward_raw.last_counter` (device int32[2] = {R, overflow}).
    """
    lib = _lib.load()
    if means3D.ndim != 2 or means3D.shape[1] != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    dev = means3D.device
    if not means3D.is_cuda:
        raise _lib.CurveGSError("means3D must be a CUDA tensor; there is no CPU path")
    P = means3D.shape[0]
    H, W = int(rs.image_height), int(rs.image_width)
    if P == 0:
        # reference: forward is skipped and the zero-filled outputs are returned (rasterize_points.cu:71-91)
        z = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=dev)
        e = torch.empty(0, dtype=torch.uint8, device=dev)
        return 0, z(1, H, W), torch.zeros(0, dtype=torch.int32, device=dev), e, e, e, z(1, H, W), z(4, H, W)
    keep: list = []
    s = _settings_struct(rs, keep)
    stream = _lib.stream(dev)

    color = torch.empty((1, H, W), dtype=torch.float32, device=dev)
    invdepth = torch.empty((1, H, W), dtype=torch.float32, device=dev)
    out_all_map = torch.zeros((4, H, W), dtype=torch.float32, device=dev) if not rs.render_geo else \
        torch.empty((4, H, W), dtype=torch.float32, device=dev)
    radii = torch.empty((P,), dtype=torch.int32, device=dev)
    geom = _bytes(lib.cg_raster_geom_bytes(P), dev)
    img = _bytes(lib.cg_raster_img_bytes(W, H), dev)

    means3D = _f32c(means3D)
    opacities = _f32c(opacities)
    scales = _f32c(scales) if scales is not None and scales.numel() else None
    rotations = _f32c(rotations) if rotations is not None and rotations.numel() else None
    cov3D = _f32c(cov3Ds_precomp) if cov3Ds_precomp is not None and cov3Ds_precomp.numel() else None
    colors = _f32c(colors_precomp) if colors_precomp is not None and colors_precomp.numel() else None
    amap = _f32c(all_maps) if all_maps is not None and all_maps.numel() else None
    if P > 0 and colors is None:
        raise _lib.CurveGSError("colors_precomp is required (the SH path is not part of the curve pipeline)")

    key = (P, W, H)
    if _policy is not None and key not in _policy._sticky and not _capturing():
        _policy._sticky[key] = torch.zeros(2, dtype=torch.int32, device=dev)
    if capacity is None and _policy is not None:
        capacity = _policy.capacity(key)
        if capacity is None and _capturing():
            raise _lib.CurveGSError(f"no binning capacity known for (P, W, H) = {key}: run the step once outside "
                                    "the CUDA graph capture (or CapacityBinning.learn(key, R)) first")
    if capacity is not None:
        cap = int(capacity)
        bin_keep = _bytes(lib.cg_raster_bin_keep_bytes(cap), dev)
        bin_scratch = _bytes(lib.cg_raster_bin_scratch_bytes(P, cap), dev)
        counter = torch.empty(2, dtype=torch.int32, device=dev)
        with _lib.on_device(dev):
            _lib.check(lib.cg_raster_fwd_capacity(
                C.byref(s), P, cap, _lib.ptr(means3D), _lib.ptr(opacities), _lib.ptr(scales), _lib.ptr(rotations),
                _lib.ptr(cov3D), _lib.ptr(colors), _lib.ptr(amap), _lib.ptr(radii), geom.data_ptr(), geom.numel(),
                img.data_ptr(), bin_keep.data_ptr(), bin_scratch.data_ptr(), color.data_ptr(), invdepth.data_ptr(),
                out_all_map.data_ptr(), counter.data_ptr(), stream), "cg_raster_fwd_capacity")
        if _policy is not None:
            _policy.observe(key, counter, cap)
        rasterize_forward_raw.last_scratch = bin_scratch
        rasterize_forward_raw.last_R = cap
        rasterize_forward_raw.last_counter = counter
        return cap, color, radii, geom, bin_keep, img, invdepth, out_all_map

    R = C.c_int64(0)
    with _lib.on_device(dev):
        _lib.check(lib.cg_raster_fwd_geom(C.byref(s), P, _lib.ptr(means3D), _lib.ptr(opacities), _lib.ptr(scales),
                                          _lib.ptr(rotations), _lib.ptr(cov3D), _lib.ptr(colors), _lib.ptr(amap),
                                          _lib.ptr(radii), geom.data_ptr(), geom.numel(), C.byref(R), stream),
                   "cg_raster_fwd_geom")
        R = int(R.value)
        Ra = _alloc_instances(R)     # (the layout inside the buffers follows R; only their size is rounded up)
        bin_keep = _bytes(lib.cg_raster_bin_keep_bytes(Ra), dev)
        bin_scratch = _bytes(lib.cg_raster_bin_scratch_bytes(P, Ra), dev)
        _lib.check(lib.cg_raster_fwd_blend(C.byref(s), P, R, geom.data_ptr(),
                                           img.data_ptr(), bin_keep.data_ptr(), bin_scratch.data_ptr(),
                                           color.data_ptr(), invdepth.data_ptr(), out_all_map.data_ptr(), stream),
                   "cg_raster_fwd_blend")
    # keep scratch reachable for debug_fetch of the sorted keys
    rasterize_forward_raw.last_scratch = bin_scratch
    rasterize_forward_raw.last_R = R
    rasterize_forward_raw.last_counter = None
    if _policy is not None:
        _policy.learn(key, R, dev)
    return R, color, radii, geom, bin_keep, img, invdepth, out_all_map


def rasterize_backward_raw(rs, means3D, radii, colors_precomp, all_maps, opacities, scales, rotations,
                           cov3Ds_precomp, grad_color, grad_invdepth, grad_all_map, geom, R, bin_keep, img,
                           need_cov3D=True, need_all_map=True):
    """The `_C.rasterize_gaussians_backward` equivalent (rasterize_points.cu:133-240).

    Returns (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales,
    dL_drotations, dL_dall_map). With need_cov3D / need_all_map False the corresponding entry is None and
    is neither allocated nor written (the autograd wrapper asks only for what autograd will use).
    """
    lib = _lib.load()
    dev = means3D.device
    P = means3D.shape[0]
    f = dict(dtype=torch.float32, device=dev)
    if P == 0:
        z = lambda *shape: torch.zeros(shape, **f)
        return z(0, 3), z(0, 1), z(0, 1), z(0, 3), z(0, 6), z(0, 0, 3), z(0, 3), z(0, 4), z(0, 4)
    keep: list = []
    s = _settings_struct(rs, keep)
    stream = _lib.stream(dev)
    d_means2D = torch.empty((P, 3), **f)
    d_colors = torch.empty((P, 1), **f)
    d_opacity = torch.empty((P, 1), **f)
    d_means3D = torch.empty((P, 3), **f)
    d_cov3D = torch.empty((P, 6), **f) if need_cov3D else None
    d_sh = torch.zeros((P, 0, 3), **f)
    d_scales = torch.empty((P, 3), **f)
    d_rot = torch.empty((P, 4), **f)
    d_all_map = torch.empty((P, 4), **f) if need_all_map else None
    scratch = _bytes(lib.cg_raster_bwd_scratch_bytes(P), dev)

    means3D = _f32c(means3D)
    opacities = _f32c(opacities)
    scales = _f32c(scales) if scales is not None and scales.numel() else None
    rotations = _f32c(rotations) if rotations is not None and rotations.numel() else None
    cov3D = _f32c(cov3Ds_precomp) if cov3Ds_precomp is not None and cov3Ds_precomp.numel() else None
    if cov3D is not None:
        d_scales.zero_()
        d_rot.zero_()
    g_color = _f32c(grad_color)
    g_invd = _f32c(grad_invdepth) if grad_invdepth is not None and grad_invdepth.numel() else None
    g_map = _f32c(grad_all_map) if grad_all_map is not None and grad_all_map.numel() else None
    with _lib.on_device(dev):
        _lib.check(lib.cg_raster_bwd(C.byref(s), P, int(R), _lib.ptr(means3D), _lib.ptr(opacities), _lib.ptr(scales),
                                     _lib.ptr(rotations), _lib.ptr(cov3D), _lib.ptr(radii), geom.data_ptr(),
                                     img.data_ptr(), bin_keep.data_ptr(), g_color.data_ptr(), _lib.ptr(g_invd),
                                     _lib.ptr(g_map), scratch.data_ptr(), d_means2D.data_ptr(), d_colors.data_ptr(),
                                     d_opacity.data_ptr(), d_means3D.data_ptr(), _lib.ptr(d_cov3D),
                                     d_scales.data_ptr(), d_rot.data_ptr(), _lib.ptr(d_all_map), stream),
                   "cg_raster_bwd")
    return d_means2D, d_colors, d_opacity, d_means3D, d_cov3D, d_sh, d_scales, d_rot, d_all_map


def mark_visible(positions, viewmatrix, projmatrix):
    lib = _lib.load()
    P = positions.shape[0]
    present = torch.zeros((P,), dtype=torch.bool, device=positions.device)
    if P:
        positions = _f32c(positions)
        vm, pm = _f32c(viewmatrix), _f32c(projmatrix)
        with _lib.on_device(positions.device):
            _lib.check(lib.cg_mark_visible(P, positions.data_ptr(), vm.data_ptr(), pm.data_ptr(), present.data_ptr(),
                                           _lib.stream(positions.device)), "cg_mark_visible")
    return present


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, all_map,
                        raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                     cov3Ds_precomp, all_map, raster_settings)


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, all_maps,
                raster_settings):
        if sh is not None and sh.numel() != 0:
            raise _lib.CurveGSError("SH colours are not part of the curve-Gaussian hot path; pass colors_precomp")
        (num_rendered, color, radii, geom, bin_keep, img, invdepths, out_all_map) = rasterize_forward_raw(
            raster_settings, means3D, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, all_maps)
        ctx.raster_settings = raster_settings
        ctx.num_rendered = num_rendered
        # unused outputs then arrive as None instead of materialised zero images; the
        # kernels skip those channels, which is bit-identical to adding zeros
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(radii)
        ctx.save_for_backward(colors_precomp, all_maps, means3D, scales, rotations, cov3Ds_precomp, radii, opacities,
                              geom, bin_keep, img)
        return color, radii, invdepths, out_all_map

    @staticmethod
    def backward(ctx, grad_out_color, _, grad_out_depth, grad_out_all_map):
        rs = ctx.raster_settings
        (colors_precomp, all_maps, means3D, scales, rotations, cov3Ds_precomp, radii, opacities, geom, bin_keep,
         img) = ctx.saved_tensors
        if grad_out_color is None:
            grad_out_color = torch.zeros((1, rs.image_height, rs.image_width), dtype=torch.float32,
                                         device=means3D.device)
        (g_means2D, g_colors, g_opac, g_means3D, g_cov3D, g_sh, g_scales, g_rot, g_all_map) = rasterize_backward_raw(
            rs, means3D, radii, colors_precomp, all_maps, opacities, scales, rotations, cov3Ds_precomp,
            grad_out_color, grad_out_depth, grad_out_all_map, geom, ctx.num_rendered, bin_keep, img,
            need_cov3D=cov3Ds_precomp is not None and cov3Ds_precomp.numel() > 0,
            # no upstream gradient on the all_map image -> dL/dall_map is identically zero: hand autograd None
            need_all_map=grad_out_all_map is not None and all_maps is not None and all_maps.numel() > 0)
        if opacities.dim() == 1:
            g_opac = g_opac.view(-1)

        def like(g, t):
            if g is None or t is None or t.numel() == 0:
                return None
            return g.view(t.shape) if g.numel() == t.numel() else g

        return (g_means3D, g_means2D, None, like(g_colors, colors_precomp), g_opac, like(g_scales, scales),
                like(g_rot, rotations), like(g_cov3D, cov3Ds_precomp), like(g_all_map, all_maps), None)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        with torch.no_grad():
            rs = self.raster_settings
            return mark_visible(positions, rs.viewmatrix, rs.projmatrix)

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None, all_map=None):
        rs = self.raster_settings
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        empty = torch.Tensor([])
        if shs is None:
            shs = empty
        if colors_precomp is None:
            colors_precomp = empty
        if scales is None:
            scales = empty
        if rotations is None:
            rotations = empty
        if cov3D_precomp is None:
            cov3D_precomp = empty
        if all_map is None:
            all_map = empty
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations,
                                   cov3D_precomp, all_map, rs)
