"""GPU parity: libcurvegs rasterizer vs the UNMODIFIED reference CUDA rasterizer
(oracle/_ref/diff_cur_rasterization_C.so, built from /root/reference by
oracle/build_ref.sh) on identical seeded inputs.

Bars (BASELINE.json north_star): bit-exact radii / tile counts / sort keys /
point list / tile ranges / n_contrib; <= 1e-5 max-rel on pixels and gradients.
"""
import math

import numpy as np
import pytest
import torch

from tests import refload
from curve_gaussian_b200 import _lib, synth
from curve_gaussian_b200.rasterizer import (GaussianRasterizationSettings, rasterize_backward_raw,
                                            rasterize_forward_raw)

pytestmark = pytest.mark.gpu

PIX_TOL = 1e-5
GRAD_TOL = 1e-5


def align128(x):
    return (x + 127) & ~127


def decode_ref_buffers(geom, binning, img, P, R, N):
    """Typed views into the reference's opaque byte buffers (rasterizer_impl.cu:155-194)."""
    out = {}
    base = geom.data_ptr()
    off = align128(base) - base

    def take(buf, off, dtype, count, itemsize):
        off = align128(buf.data_ptr() + off) - buf.data_ptr()
        view = buf[off:off + count * itemsize].view(dtype)
        return view, off + count * itemsize

    out["depths"], off = take(geom, 0, torch.float32, P, 4)
    _, off = take(geom, off, torch.uint8, 3 * P, 1)
    _, off = take(geom, off, torch.int32, P, 4)
    out["means2D"], off = take(geom, off, torch.float32, 2 * P, 4)
    _, off = take(geom, off, torch.float32, 6 * P, 4)
    out["conic_opacity"], off = take(geom, off, torch.float32, 4 * P, 4)
    _, off = take(geom, off, torch.float32, P, 4)
    out["tiles_touched"], off = take(geom, off, torch.int32, P, 4)
    out["accum_alpha"], off = take(img, 0, torch.float32, N, 4)
    out["n_contrib"], off = take(img, off, torch.int32, N, 4)
    out["ranges"], off = take(img, off, torch.int32, 2 * N, 4)
    if R > 0:
        out["point_list"], off = take(binning, 0, torch.int32, R, 4)
        _, off = take(binning, off, torch.int32, R, 4)
        out["keys"], off = take(binning, off, torch.int64, R, 8)
    return out


def fetch(which, P, R, W, H, geom, img, bin_keep, scratch, dtype, count):
    lib = _lib.load()
    dst = torch.empty(count, dtype=dtype, device=geom.device)
    _lib.check(lib.cg_raster_debug_fetch(which, P, R, W, H, geom.data_ptr(), img.data_ptr(), bin_keep.data_ptr(),
                                         scratch.data_ptr(), dst.data_ptr(),
                                         torch.cuda.current_stream().cuda_stream), "debug_fetch")
    torch.cuda.synchronize()
    return dst


def max_rel(a, b):
    a = a.double().flatten()
    b = b.double().flatten()
    denom = b.abs().max().item()
    if denom == 0:
        return (a - b).abs().max().item()
    return ((a - b).abs().max() / denom).item()


def make_case(name, dev):
    if name == "cloud_small":
        W, H, P = 250, 190, 3000
        means, scales, rots, opac, colors, amap = synth.random_gaussians(P, seed=1)
    elif name == "cloud_dense":
        W, H, P = 320, 240, 20000
        means, scales, rots, opac, colors, amap = synth.random_gaussians(P, seed=2, scale_lo=0.002, scale_hi=0.02)
    elif name == "discs":
        # thin discs like sampled curve Gaussians: scale (tiny, w, w), non-unit quaternions
        W, H, P = 400, 300, 30000
        means, scales, rots, opac, colors, amap = synth.random_gaussians(P, seed=3)
        scales = torch.stack([torch.full((P,), 7e-4), torch.full((P,), 5e-3), torch.full((P,), 5e-3)], 1)
        rots = rots * (0.7 + 0.6 * torch.rand(P, 1, generator=torch.Generator().manual_seed(5)))
        opac = torch.full((P, 1), 0.6)
        colors = torch.ones(P, 1)
    elif name == "behind_and_offscreen":
        W, H, P = 128, 128, 4000
        means, scales, rots, opac, colors, amap = synth.random_gaussians(P, seed=4)
        means = means * 6.0 - 2.5   # many behind the camera / far off screen
    elif name == "ties_and_extremes":
        # equal sort keys (exact duplicates and a whole plane of Gaussians at one view-space depth), screen-filling
        # and sub-pixel splats, zero / tiny / saturated opacities, an image that is not a multiple of the tile size
        W, H, P = 203, 117, 6000
        means, scales, rots, opac, colors, amap = synth.random_gaussians(P, seed=6)
        cam0 = synth.random_cameras(1, W, H, seed=15)[0]
        g = torch.Generator().manual_seed(9)
        # a plane of constant view-space depth: p = c + a*right + b*up + d*forward with the camera's own axes
        R_w2c = cam0.world_view_transform[:3, :3]          # columns: camera axes in world coordinates
        centre = cam0.camera_center
        ab = (torch.rand(2000, 2, generator=g) - 0.5) * 1.6
        means[:2000] = centre + ab[:, :1] * R_w2c[:, 0] + ab[:, 1:] * R_w2c[:, 1] + 2.0 * R_w2c[:, 2]
        means[2000:2500] = means[1500:2000]                 # exact duplicates: ties broken by index only
        scales[2500:2510] = 3.0                              # screen-filling
        scales[2510:2600] = 1e-6                             # far below a pixel
        opac[2600:2700] = 0.0
        opac[2700:2800] = 1.0 / 255.0
        opac[2800:2900] = 1.0
        opac[2900:2950] = 0.003
    else:
        raise KeyError(name)
    seed = {"cloud_small": 11, "cloud_dense": 12, "discs": 13, "behind_and_offscreen": 14, "ties_and_extremes": 15}[name]
    cam = synth.random_cameras(1, W, H, seed=seed)[0].to(dev)
    t = lambda x: x.to(dev).contiguous()
    return cam, t(means), t(scales), t(rots), t(opac), t(colors), t(amap)


def settings_for(cam, dev, render_geo=True, bg_val=0.0):
    return GaussianRasterizationSettings(
        image_height=cam.image_height, image_width=cam.image_width,
        tanfovx=math.tan(cam.FoVx * 0.5), tanfovy=math.tan(cam.FoVy * 0.5),
        bg=torch.full((3,), bg_val, device=dev), scale_modifier=1.0,
        viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform,
        sh_degree=0, campos=cam.camera_center, prefiltered=False, debug=False,
        antialiasing=False, render_geo=render_geo)


def run_reference(rs, means, colors, opac, scales, rots, amap, grads):
    ref = refload.ref_rasterizer()
    empty = torch.Tensor([])
    out = ref.rasterize_gaussians(rs.bg, means, colors, opac, scales, rots, rs.scale_modifier, empty, amap,
                                  rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.image_height,
                                  rs.image_width, empty, 0, rs.campos, False, rs.antialiasing, rs.render_geo, False)
    R, color, radii, geomB, binB, imgB, invd, omap = out
    g_color, g_invd, g_map = grads
    bw = ref.rasterize_gaussians_backward(rs.bg, omap, means, radii, colors, amap, opac, scales, rots,
                                          rs.scale_modifier, empty, rs.viewmatrix, rs.projmatrix, rs.tanfovx,
                                          rs.tanfovy, g_color, g_invd, g_map, empty, 0, rs.campos, geomB, R, binB,
                                          imgB, rs.antialiasing, rs.render_geo, False)
    torch.cuda.synchronize()
    return out, bw


@pytest.mark.parametrize("case", ["cloud_small", "cloud_dense", "discs", "behind_and_offscreen", "ties_and_extremes"])
@pytest.mark.parametrize("bg_val", [0.0, 0.3])
def test_forward_backward_match_reference(cuda_dev, case, bg_val):
    if refload.ref_rasterizer() is None:
        pytest.skip("oracle/_ref/diff_cur_rasterization_C.so not built")
    dev = cuda_dev
    cam, means, scales, rots, opac, colors, amap = make_case(case, dev)
    rs = settings_for(cam, dev, True, bg_val)
    H, W = rs.image_height, rs.image_width
    P = means.shape[0]
    gen = torch.Generator(device="cpu").manual_seed(7)
    g_color = torch.randn(1, H, W, generator=gen).to(dev)
    g_invd = torch.randn(1, H, W, generator=gen).to(dev) * 0.1
    g_map = torch.randn(4, H, W, generator=gen).to(dev) * 0.1

    (R_ref, color_ref, radii_ref, geomB, binB, imgB, invd_ref, omap_ref), bw_ref = run_reference(
        rs, means, colors, opac, scales, rots, amap, (g_color, g_invd, g_map))

    R, color, radii, geom, bin_keep, img, invd, omap = rasterize_forward_raw(
        rs, means, colors, opac, scales, rots, None, amap)
    scratch = rasterize_forward_raw.last_scratch
    torch.cuda.synchronize()

    # ---- integer state: bit-exact
    assert R == R_ref
    assert torch.equal(radii, radii_ref)
    dec = decode_ref_buffers(geomB, binB, imgB, P, R_ref, W * H)
    vis = radii_ref > 0
    tiles = fetch(3, P, R, W, H, geom, img, bin_keep, scratch, torch.int32, P)
    assert torch.equal(tiles, dec["tiles_touched"])
    depths = fetch(5, P, R, W, H, geom, img, bin_keep, scratch, torch.float32, P)
    assert torch.equal(depths[vis].view(torch.int32), dec["depths"][vis].view(torch.int32))
    xy = fetch(4, P, R, W, H, geom, img, bin_keep, scratch, torch.float32, 2 * P).view(P, 2)
    assert torch.equal(xy[vis].view(torch.int32), dec["means2D"].view(P, 2)[vis].view(torch.int32))
    co = fetch(6, P, R, W, H, geom, img, bin_keep, scratch, torch.float32, 4 * P).view(P, 4)
    assert torch.equal(co[vis].view(torch.int32), dec["conic_opacity"].view(P, 4)[vis].view(torch.int32))
    if R > 0:
        keys = fetch(0, P, R, W, H, geom, img, bin_keep, scratch, torch.int64, R)
        assert torch.equal(keys, dec["keys"])
        plist = fetch(1, P, R, W, H, geom, img, bin_keep, scratch, torch.int32, R)
        assert torch.equal(plist, dec["point_list"])
    ntiles = ((W + 15) // 16) * ((H + 15) // 16)
    ranges = fetch(2, P, R, W, H, geom, img, bin_keep, scratch, torch.int32, 2 * ntiles)
    assert torch.equal(ranges, dec["ranges"][:2 * ntiles])
    ncon = fetch(7, P, R, W, H, geom, img, bin_keep, scratch, torch.int32, W * H)
    assert torch.equal(ncon, dec["n_contrib"])

    # ---- pixels: <= 1e-5 max-rel (in practice bit-identical)
    for name, a, b in (("color", color, color_ref), ("invdepth", invd, invd_ref), ("all_map", omap, omap_ref)):
        assert max_rel(a, b) <= PIX_TOL, name
    fT = fetch(8, P, R, W, H, geom, img, bin_keep, scratch, torch.float32, W * H)
    assert max_rel(fT, dec["accum_alpha"]) <= PIX_TOL

    # ---- gradients
    bw = rasterize_backward_raw(rs, means, radii, colors, amap, opac, scales, rots, None, g_color, g_invd, g_map,
                                geom, R, bin_keep, img)
    torch.cuda.synchronize()
    names = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales",
             "dL_drotations", "dL_dall_map"]
    # the reference's own run-to-run noise (fp32 atomics) sets the floor for this comparison
    _, bw_ref2 = run_reference(rs, means, colors, opac, scales, rots, amap, (g_color, g_invd, g_map))
    for n, a, b, b2 in zip(names, bw, bw_ref, bw_ref2):
        if n == "dL_dsh":
            continue
        noise = max_rel(b2, b)
        err = max_rel(a, b)
        assert err <= max(GRAD_TOL, 4 * noise), f"{n}: err {err:.3e} (reference self-noise {noise:.3e})"


def test_color_only_backward_skips_unused_channels(cuda_dev):
    """None grads for invdepth/all_map (set_materialize_grads(False)) == reference fed with zeros."""
    if refload.ref_rasterizer() is None:
        pytest.skip("oracle/_ref/diff_cur_rasterization_C.so not built")
    dev = cuda_dev
    cam, means, scales, rots, opac, colors, amap = make_case("discs", dev)
    rs = settings_for(cam, dev, True, 0.0)
    H, W = rs.image_height, rs.image_width
    g_color = torch.randn(1, H, W, generator=torch.Generator().manual_seed(3)).to(dev)
    zeros1 = torch.zeros(1, H, W, device=dev)
    zeros4 = torch.zeros(4, H, W, device=dev)
    _, bw_ref = run_reference(rs, means, colors, opac, scales, rots, amap, (g_color, zeros1, zeros4))
    R, color, radii, geom, bin_keep, img, invd, omap = rasterize_forward_raw(rs, means, colors, opac, scales, rots,
                                                                              None, amap)
    bw = rasterize_backward_raw(rs, means, radii, colors, amap, opac, scales, rots, None, g_color, None, None,
                                geom, R, bin_keep, img)
    torch.cuda.synchronize()
    for i in (0, 1, 2, 3, 6, 7, 8):
        assert max_rel(bw[i], bw_ref[i]) <= GRAD_TOL, i
