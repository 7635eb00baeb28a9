"""GPU, BASELINE.json's full size (config C4: 10k cubic Beziers x 100 samples = 1M curve-Gaussians, 1920x1080):
(a) the same bit-exact / 1e-5 bars against the UNMODIFIED reference CUDA rasterizer as the small cases, and
(b) size-independent properties that need no oracle: sortedness of the rebuilt 64-bit keys, the point list being
the multiset of emitted instances, tile ranges partitioning [0, R), forward determinism, and linearity of the
backward pass in the upstream gradient."""
import math

import pytest
import torch

from tests import parity, refload
from tests.test_gpu_raster_vs_reference import decode_ref_buffers, fetch, max_rel, run_reference, settings_for
from curve_gaussian_b200 import synth
from curve_gaussian_b200.activation import curve_activate
from curve_gaussian_b200.curve_model import GaussianCurveModel
from curve_gaussian_b200.rasterizer import rasterize_backward_raw, rasterize_forward_raw

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c4(cuda_dev):
    dev = cuda_dev
    B, n, W, H = 10000, 100, 1920, 1080
    cp, width, opl, isb = synth.random_curves(B, seed=0)
    model = GaussianCurveModel(0, n_gaussians=n, device=dev).create_from_curves(cp, width, opl, isb)
    cam = synth.random_cameras(1, W, H, seed=0)[0].to(dev)
    with torch.no_grad():
        rots, opac, scales, amap = curve_activate(model._xyz, model._rotation, model._scaling, model._opacity,
                                                  model._mask, n, cam.camera_center, cam.world_view_transform)
        means = model._xyz.detach().contiguous()
    colors = torch.ones(B * n, 1, device=dev)
    rs = settings_for(cam, dev, True, 0.0)
    g_color = torch.randn(1, H, W, generator=torch.Generator().manual_seed(5)).to(dev)
    return dict(rs=rs, means=means, scales=scales.contiguous(), rots=rots.contiguous(), opac=opac.contiguous(),
                colors=colors, amap=amap.contiguous(), g_color=g_color, P=B * n, W=W, H=H)


def test_c4_properties(c4):
    rs, P, W, H = c4["rs"], c4["P"], c4["W"], c4["H"]
    args = (rs, c4["means"], c4["colors"], c4["opac"], c4["scales"], c4["rots"], None, c4["amap"])
    R, color, radii, geom, bin_keep, img, invd, omap = rasterize_forward_raw(*args)
    scratch = rasterize_forward_raw.last_scratch
    assert R > 1_000_000
    keys = fetch(0, P, R, W, H, geom, img, bin_keep, scratch, torch.int64, R)
    plist = fetch(1, P, R, W, H, geom, img, bin_keep, scratch, torch.int32, R)
    tiles = fetch(3, P, R, W, H, geom, img, bin_keep, scratch, torch.int32, P)
    assert bool((keys[1:] >= keys[:-1]).all()), "keys sorted by (tile, depth)"
    # stable: equal keys keep ascending Gaussian index
    same = keys[1:] == keys[:-1]
    assert bool((plist[1:][same] > plist[:-1][same]).all())
    # the point list holds every Gaussian exactly tiles_touched times
    assert int(tiles.sum()) == R
    assert torch.equal(torch.bincount(plist.long(), minlength=P).int(), tiles)
    # ranges partition [0, R) in tile order
    ntiles = ((W + 15) // 16) * ((H + 15) // 16)
    ranges = fetch(2, P, R, W, H, geom, img, bin_keep, scratch, torch.int32, 2 * ntiles).view(ntiles, 2).long()
    ne = ranges[ranges[:, 1] > ranges[:, 0]]
    assert int(ne[0, 0]) == 0 and int(ne[-1, 1]) == R and torch.equal(ne[1:, 0], ne[:-1, 1])
    tile_of = (keys >> 32)
    assert torch.equal(tile_of[ne[:, 0]], torch.nonzero(ranges[:, 1] > ranges[:, 0]).flatten())
    # forward is deterministic, bit for bit
    R2, color2, radii2, *_rest = rasterize_forward_raw(*args)
    assert R2 == R and torch.equal(color, color2) and torch.equal(radii, radii2)
    # backward is linear in the upstream gradient (fp32 atomics: tolerance, not equality)
    g = c4["g_color"]
    bw1 = rasterize_backward_raw(rs, c4["means"], radii, c4["colors"], c4["amap"], c4["opac"], c4["scales"], c4["rots"],
                                 None, g, None, None, geom, R, bin_keep, img)
    bw2 = rasterize_backward_raw(rs, c4["means"], radii, c4["colors"], c4["amap"], c4["opac"], c4["scales"], c4["rots"],
                                 None, 2.0 * g, None, None, geom, R, bin_keep, img)
    for i in (0, 2, 3, 6, 7):
        assert max_rel(bw2[i], 2.0 * bw1[i]) <= 1e-5, i


def test_c4_matches_reference_cuda(c4):
    if refload.ref_rasterizer() is None:
        pytest.skip("oracle/_ref/diff_cur_rasterization_C.so not built")
    rs, P, W, H = c4["rs"], c4["P"], c4["W"], c4["H"]
    dev = c4["means"].device
    g_color = c4["g_color"]
    z1, z4 = torch.zeros(1, H, W, device=dev), torch.zeros(4, H, W, device=dev)
    (R_ref, color_ref, radii_ref, geomB, binB, imgB, invd_ref, omap_ref), bw_ref = run_reference(
        rs, c4["means"], c4["colors"], c4["opac"], c4["scales"], c4["rots"], c4["amap"], (g_color, z1, z4))
    R, color, radii, geom, bin_keep, img, invd, omap = rasterize_forward_raw(
        rs, c4["means"], c4["colors"], c4["opac"], c4["scales"], c4["rots"], None, c4["amap"])
    scratch = rasterize_forward_raw.last_scratch
    assert R == R_ref and torch.equal(radii, radii_ref)
    dec = decode_ref_buffers(geomB, binB, imgB, P, R_ref, W * H)
    assert torch.equal(fetch(0, P, R, W, H, geom, img, bin_keep, scratch, torch.int64, R), dec["keys"])
    assert torch.equal(fetch(1, P, R, W, H, geom, img, bin_keep, scratch, torch.int32, R), dec["point_list"])
    assert torch.equal(fetch(7, P, R, W, H, geom, img, bin_keep, scratch, torch.int32, W * H), dec["n_contrib"])
    # pixels: the bar is 1e-5 max-rel; they are in fact bit-identical
    assert torch.equal(color, color_ref) and torch.equal(invd, invd_ref) and torch.equal(omap, omap_ref)
    bw = rasterize_backward_raw(rs, c4["means"], radii, c4["colors"], c4["amap"], c4["opac"], c4["scales"], c4["rots"],
                                None, g_color, None, None, geom, R, bin_keep, img)
    _, bw_ref2 = run_reference(rs, c4["means"], c4["colors"], c4["opac"], c4["scales"], c4["rots"], c4["amap"],
                               (g_color, z1, z4))
    _, bw_ref3 = run_reference(rs, c4["means"], c4["colors"], c4["opac"], c4["scales"], c4["rots"], c4["amap"],
                               (g_color, z1, z4))
    for i, name in ((0, "dL_dmeans2D"), (1, "dL_dcolors"), (2, "dL_dopacity"), (3, "dL_dmeans3D"), (6, "dL_dscales"),
                    (7, "dL_drotations")):
        parity.check("raster_vs_reference", "C4_colour_only", name, bw[i], bw_ref[i], [bw_ref2[i], bw_ref3[i]])


def test_c4_all_upstream_gradients_match_reference_cuda(c4):
    """C4 with dL/dcolour, dL/dinvdepth and dL/dall_map all non-zero: the 16-term geometry variant of the backward
    (blend_bwd<1,1>) at full size, every returned gradient against the reference."""
    if refload.ref_rasterizer() is None:
        pytest.skip("oracle/_ref/diff_cur_rasterization_C.so not built")
    rs, P, W, H = c4["rs"], c4["P"], c4["W"], c4["H"]
    dev = c4["means"].device
    gen = torch.Generator().manual_seed(11)
    g_color = c4["g_color"]
    g_invd = (torch.randn(1, H, W, generator=gen) * 0.1).to(dev)
    g_map = (torch.randn(4, H, W, generator=gen) * 0.1).to(dev)
    refs = [run_reference(rs, c4["means"], c4["colors"], c4["opac"], c4["scales"], c4["rots"], c4["amap"],
                          (g_color, g_invd, g_map))[1] for _ in range(3)]
    R, color, radii, geom, bin_keep, img, invd, omap = rasterize_forward_raw(
        rs, c4["means"], c4["colors"], c4["opac"], c4["scales"], c4["rots"], None, c4["amap"])
    bw = rasterize_backward_raw(rs, c4["means"], radii, c4["colors"], c4["amap"], c4["opac"], c4["scales"], c4["rots"],
                                None, g_color, g_invd, g_map, geom, R, bin_keep, img)
    torch.cuda.synchronize()
    names = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales",
             "dL_drotations", "dL_dall_map"]
    for i, name in enumerate(names):
        if name == "dL_dsh":
            continue
        parity.check("raster_vs_reference", "C4_all_upstream", name, bw[i], refs[0][i], [refs[1][i], refs[2][i]])
