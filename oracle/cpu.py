"""TEST INFRASTRUCTURE ONLY — numpy front-end of the C oracle (oracle/*.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module; the product path never does.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
_lib = None

_f = np.float32


def build(force: bool = False) -> str:
    srcs = [os.path.join(HERE, n) for n in ("raster_oracle.c", "ssim_oracle.c", "knn_oracle.c", "Makefile")]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        r = subprocess.run(["make", "-C", HERE, "-B"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    return LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.or_bin.restype = C.c_int64
        _lib.or_higher_msb.restype = C.c_uint32
        _lib.or_key_tile_bits.restype = C.c_uint32
        _lib.or_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dtype=_f):
    return None if a is None else np.ascontiguousarray(a, dtype=dtype)


def num_threads() -> int:
    return int(lib().or_num_threads())


class Geom:
    pass


def preprocess(means, scales, rots, opac, vm, pm, W, H, tanx, tany, mod=1.0, antialiasing=False, cov3D=None):
    P = means.shape[0]
    means, scales, rots, opac, vm, pm, cov3D = map(_c, (means, scales, rots, opac, vm, pm, cov3D))
    g = Geom()
    g.P, g.W, g.H = P, W, H
    g.radii = np.zeros(P, np.int32)
    g.xy = np.zeros((P, 2), _f)
    g.depth = np.zeros(P, _f)
    g.conic_o = np.zeros((P, 4), _f)
    g.tiles = np.zeros(P, np.uint32)
    g.rect = np.zeros((P, 4), np.int32)
    g.cov3D = np.zeros((P, 6), _f)
    lib().or_preprocess(C.c_int64(P), _p(means), _p(scales), _p(rots), _p(opac), _p(cov3D), C.c_float(mod), _p(vm),
                        _p(pm), C.c_int(W), C.c_int(H), C.c_float(tanx), C.c_float(tany), C.c_int(int(antialiasing)),
                        _p(g.radii), _p(g.xy), _p(g.depth), _p(g.conic_o), _p(g.tiles), _p(g.rect), _p(g.cov3D))
    return g


def bin_tiles(g: Geom):
    R = int(g.tiles.astype(np.int64).sum())
    ntiles = ((g.W + 15) // 16) * ((g.H + 15) // 16)
    keys = np.zeros(max(R, 1), np.uint64)
    vals = np.zeros(max(R, 1), np.uint32)
    ranges = np.zeros((ntiles, 2), np.uint32)
    R2 = lib().or_bin(C.c_int64(g.P), _p(g.tiles), _p(g.rect), _p(g.depth), C.c_int(g.W), C.c_int(g.H), _p(keys),
                      _p(vals), _p(ranges))
    assert R2 == R
    return keys[:R], vals[:R], ranges


def bin_from_arrays(tiles, rect, depth, W, H):
    g = Geom()
    g.P, g.W, g.H = tiles.shape[0], W, H
    g.tiles, g.rect, g.depth = _c(tiles, np.uint32), _c(rect, np.int32), _c(depth)
    return bin_tiles(g)


def blend_fwd(g: Geom, ranges, point_list, colors, all_map, bg, render_geo=True):
    W, H = g.W, g.H
    out = Geom()
    out.color = np.zeros((1, H, W), _f)
    out.invdepth = np.zeros((1, H, W), _f)
    out.all_map = np.zeros((4, H, W), _f)
    out.final_T = np.zeros(H * W, _f)
    out.n_contrib = np.zeros(H * W, np.uint32)
    colors, all_map, bg = _c(colors), _c(all_map), _c(bg)
    ranges = _c(ranges, np.uint32)
    point_list = _c(point_list, np.uint32)
    lib().or_blend_fwd(C.c_int(W), C.c_int(H), _p(ranges), _p(point_list), _p(g.xy), _p(g.conic_o), _p(colors),
                       _p(g.depth), _p(all_map), C.c_int(int(render_geo)), _p(bg), _p(out.color), _p(out.invdepth),
                       _p(out.all_map), _p(out.final_T), _p(out.n_contrib))
    return out


def blend_bwd(g: Geom, ranges, point_list, colors, all_map, bg, fwd, dL_dcolor, dL_dinvd=None, dL_dmap=None,
              render_geo=True):
    W, H, P = g.W, g.H, g.P
    acc = Geom()
    acc.d_mean2D = np.zeros((P, 3), _f)
    acc.d_conic = np.zeros((P, 4), _f)
    acc.d_opacity = np.zeros(P, _f)
    acc.d_colors = np.zeros(P, _f)
    acc.d_invdepths = np.zeros(P, _f)
    acc.d_all_map = np.zeros((P, 4), _f)
    colors, all_map, bg = _c(colors), _c(all_map), _c(bg)
    ranges = _c(ranges, np.uint32)
    point_list = _c(point_list, np.uint32)
    dL_dcolor, dL_dinvd, dL_dmap = _c(dL_dcolor), _c(dL_dinvd), _c(dL_dmap)
    lib().or_blend_bwd(C.c_int(W), C.c_int(H), _p(ranges), _p(point_list), _p(g.xy), _p(g.conic_o), _p(colors),
                       _p(g.depth), _p(all_map), C.c_int(int(render_geo)), _p(bg), _p(fwd.final_T),
                       _p(fwd.n_contrib), _p(dL_dcolor), _p(dL_dinvd), _p(dL_dmap), _p(acc.d_mean2D),
                       _p(acc.d_conic), _p(acc.d_opacity), _p(acc.d_colors), _p(acc.d_invdepths), _p(acc.d_all_map))
    acc.has_invd = dL_dinvd is not None
    return acc


def preprocess_bwd(g: Geom, acc, means, scales, rots, opac, vm, pm, tanx, tany, mod=1.0, antialiasing=False,
                   cov3D=None):
    P = g.P
    means, scales, rots, opac, vm, pm, cov3D = map(_c, (means, scales, rots, opac, vm, pm, cov3D))
    out = Geom()
    out.d_means3D = np.zeros((P, 3), _f)
    out.d_cov3D = np.zeros((P, 6), _f)
    out.d_scales = np.zeros((P, 3), _f)
    out.d_rots = np.zeros((P, 4), _f)
    out.d_opacity = acc.d_opacity.copy()
    lib().or_preprocess_bwd(C.c_int64(P), _p(means), _p(scales), _p(rots), _p(opac), _p(cov3D), C.c_float(mod),
                            _p(g.radii), _p(vm), _p(pm), C.c_int(g.W), C.c_int(g.H), C.c_float(tanx),
                            C.c_float(tany), C.c_int(int(antialiasing)), _p(acc.d_mean2D), _p(acc.d_conic),
                            _p(acc.d_invdepths) if acc.has_invd else None, _p(out.d_opacity), _p(out.d_means3D),
                            _p(out.d_cov3D), _p(out.d_scales), _p(out.d_rots))
    return out


def rasterize_fwd_bwd(means, scales, rots, opac, colors, all_map, vm, pm, campos, W, H, tanx, tany, bg, dL_dcolor=None,
                      dL_dinvd=None, dL_dmap=None, render_geo=True, mod=1.0, antialiasing=False):
    """Whole reference pipeline on the CPU; returns a dict mirroring the reference's outputs."""
    g = preprocess(means, scales, rots, np.asarray(opac).reshape(-1), vm, pm, W, H, tanx, tany, mod, antialiasing)
    keys, vals, ranges = bin_tiles(g)
    fwd = blend_fwd(g, ranges, vals, np.asarray(colors).reshape(-1), all_map, bg, render_geo)
    res = dict(R=len(keys), radii=g.radii, keys=keys, point_list=vals, ranges=ranges, color=fwd.color,
               invdepth=fwd.invdepth, all_map=fwd.all_map, final_T=fwd.final_T, n_contrib=fwd.n_contrib, geom=g)
    if dL_dcolor is not None:
        acc = blend_bwd(g, ranges, vals, np.asarray(colors).reshape(-1), all_map, bg, fwd, dL_dcolor, dL_dinvd,
                        dL_dmap, render_geo)
        pb = preprocess_bwd(g, acc, means, scales, rots, np.asarray(opac).reshape(-1), vm, pm, tanx, tany, mod,
                            antialiasing)
        res.update(dL_dmeans2D=acc.d_mean2D, dL_dcolors=acc.d_colors, dL_dopacity=pb.d_opacity,
                   dL_dmeans3D=pb.d_means3D, dL_dcov3D=pb.d_cov3D, dL_dscales=pb.d_scales, dL_drotations=pb.d_rots,
                   dL_dall_map=acc.d_all_map, dL_dconic=acc.d_conic)
    return res


def ssim_fwd(img1, img2, C1=0.01 ** 2, C2=0.03 ** 2, train=True):
    img1, img2 = _c(img1), _c(img2)
    B, CH, H, W = img1.shape
    m = np.zeros_like(img1)
    d1 = np.zeros_like(img1) if train else None
    d2 = np.zeros_like(img1) if train else None
    d3 = np.zeros_like(img1) if train else None
    lib().or_ssim_fwd(C.c_int(B), C.c_int(CH), C.c_int(H), C.c_int(W), C.c_float(C1), C.c_float(C2), _p(img1),
                      _p(img2), _p(m), _p(d1), _p(d2), _p(d3))
    return m, d1, d2, d3


def ssim_bwd(img1, img2, dL_dmap, d1, d2, d3):
    img1, img2, dL_dmap = _c(img1), _c(img2), _c(dL_dmap)
    B, CH, H, W = img1.shape
    out = np.zeros_like(img1)
    lib().or_ssim_bwd(C.c_int(B), C.c_int(CH), C.c_int(H), C.c_int(W), _p(img1), _p(img2), _p(dL_dmap), _p(d1),
                      _p(d2), _p(d3), _p(out))
    return out


def knn_mean_dist2(points):
    points = _c(points)
    out = np.zeros(points.shape[0], _f)
    lib().or_knn_mean_dist2(C.c_int64(points.shape[0]), _p(points), _p(out))
    return out
