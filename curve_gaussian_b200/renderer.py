"""Drop-in for the reference `gaussian_renderer.render` (gaussian_renderer/__init__.py:18-157):
same signature, same output dict, rasterization through libcurvegs.so."""
from __future__ import annotations

import math

import torch

from .activation import curve_activate
from .loss import rotate_channels
from .rasterizer import GaussianRasterizationSettings, GaussianRasterizer


class RenderPackage(dict):
    """The output dict of render() with the same keys as the reference
    (gaussian_renderer/__init__.py:147-155). `visibility_filter` (a host-synchronising
    nonzero()) and `rend_dir` (a per-pixel rotation nothing in train.py's loss reads) are
    produced on first access instead of on every call; every other access pattern of a
    dict (in, keys(), items(), len, iteration, get) sees them as present."""

    def __init__(self, eager, lazy):
        super().__init__(eager)
        self._lazy = dict(lazy)

    def _force(self, key):
        fn = self._lazy.pop(key, None)
        if fn is not None:
            super().__setitem__(key, fn())

    def _force_all(self):
        for k in list(self._lazy):
            self._force(k)

    def __missing__(self, key):
        if key in self._lazy:
            self._force(key)
            return super().__getitem__(key)
        raise KeyError(key)

    def __contains__(self, key):
        return super().__contains__(key) or key in self._lazy

    def get(self, key, default=None):
        return self[key] if key in self else default

    def __len__(self):
        return super().__len__() + len(self._lazy)

    def __iter__(self):
        self._force_all()
        return super().__iter__()

    def keys(self):
        self._force_all()
        return super().keys()

    def values(self):
        self._force_all()
        return super().values()

    def items(self):
        self._force_all()
        return super().items()


_ONES = {}


def _ones(P: int, dev) -> torch.Tensor:
    """(P,1) ones (the single colour channel, gaussian_renderer/__init__.py:97), cached per size and device."""
    # Entries are never evicted while they may be baked into a captured CUDA graph as `colors_precomp` (a render of
    # another model size between two replays must not hand that memory back to the allocator); the cache is
    # bounded instead: past 16 shapes new ones are simply not cached.
    key = (int(P), str(dev))
    t = _ONES.get(key)
    if t is None:
        t = torch.ones(P, 1, device=dev)
        if len(_ONES) < 16:
            _ONES[key] = t
    return t


def _can_fuse(pc) -> bool:
    """The fused activation needs the curve model's raw tensors (the reference class has them too)."""
    need = ("_xyz", "_rotation", "_scaling", "_opacity", "_mask", "n_gaussians")
    if not all(hasattr(pc, k) for k in need) or getattr(pc, "fuse_activations", True) is False:
        return False
    P = pc._xyz.shape[0]
    return pc._xyz.is_cuda and pc._opacity.numel() * pc.n_gaussians == P and pc._mask.numel() == P


def render(viewpoint_camera, pc, pipe, bg_color: torch.Tensor, scaling_modifier=1.0, separate_sh=False,
           override_color=None, use_trained_exp=False, use_mask=False, mask_thr=0.01):
    """Render the scene. Background tensor (bg_color) must be on GPU."""
    # the reference makes this a non-leaf (`zeros_like(...) + 0`, then retain_grad()); a fresh leaf gives
    # callers the same `.grad` with one fill kernel instead of two
    screenspace_points = torch.zeros_like(pc.get_xyz, dtype=pc.get_xyz.dtype, requires_grad=True,
                                          device=pc.get_xyz.device)

    tanfovx = math.tan(viewpoint_camera.FoVx * 0.5)
    tanfovy = math.tan(viewpoint_camera.FoVy * 0.5)
    raster_settings = GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height),
        image_width=int(viewpoint_camera.image_width),
        tanfovx=tanfovx,
        tanfovy=tanfovy,
        bg=bg_color,
        scale_modifier=scaling_modifier,
        viewmatrix=viewpoint_camera.world_view_transform,
        projmatrix=viewpoint_camera.full_proj_transform,
        sh_degree=pc.active_sh_degree,
        campos=viewpoint_camera.camera_center,
        prefiltered=False,
        debug=bool(getattr(pipe, "debug", False)),
        antialiasing=bool(getattr(pipe, "antialiasing", False)),
        render_geo=bool(getattr(pipe, "render_geo", True)),
    )
    rasterizer = GaussianRasterizer(raster_settings=raster_settings)

    means3D = pc.get_xyz
    means2D = screenspace_points
    dev = means3D.device
    colors_precomp = _ones(means3D.shape[0], dev)
    if _can_fuse(pc):
        # one fused kernel for normalize / sigmoid / mask straight-through / main axis / all_map
        rotations, opacity, scales, input_all_map = curve_activate(
            pc._xyz, pc._rotation, pc._scaling, pc._opacity, pc._mask, pc.n_gaussians,
            viewpoint_camera.camera_center, viewpoint_camera.world_view_transform, use_mask, mask_thr)
    else:
        opacity = pc.get_opacity
        scales = pc.get_scaling
        rotations = pc.get_rotation
        if use_mask:
            sm = torch.sigmoid(pc._mask)
            mask = ((sm > mask_thr).float() - sm).detach() + sm
            scales = pc.get_scaling * mask.view(-1, 1)
            opacity = pc.get_opacity * mask.view(-1, 1)
        global_normal = pc.get_main_axis(viewpoint_camera)
        local_normal = global_normal @ viewpoint_camera.world_view_transform[:3, :3]
        input_all_map = torch.cat([local_normal, torch.ones_like(local_normal[:, :1])], dim=1)

    rendered_image, radii, depth_image, out_all_map = rasterizer(
        means3D=means3D, means2D=means2D, shs=None, colors_precomp=colors_precomp, opacities=opacity,
        scales=scales, rotations=rotations, all_map=input_all_map, cov3D_precomp=None)

    if use_trained_exp:
        exposure = pc.get_exposure_from_name(viewpoint_camera.image_name)
        rendered_image = torch.matmul(rendered_image.permute(1, 2, 0), exposure[:3, :3]).permute(2, 0, 1) \
            + exposure[:3, 3, None, None]

    raw_image = rendered_image
    rendered_alpha = out_all_map[3:4, ]
    wvt = viewpoint_camera.world_view_transform

    return RenderPackage({
        "render_raw": raw_image,     # extra key: the render before clamp(0,1), for losses that fuse the clamp
        "viewspace_points": screenspace_points,
        "radii": radii,
        "depth": depth_image,
        "rend_alpha": rendered_alpha,
    }, {
        "render": lambda: raw_image.clamp(0, 1),
        "visibility_filter": lambda: (radii > 0).nonzero(),
        # (dir.permute(1,2,0) @ wvt[:3,:3].T).permute(2,0,1) as one per-pixel kernel
        "rend_dir": lambda: rotate_channels(out_all_map[0:3], wvt[:3, :3]),
    })
