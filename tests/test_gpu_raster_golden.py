"""GPU: libcurvegs rasterizer against the committed golden vectors of the reference CUDA
rasterizer (tests/golden/raster_*.npz) — this holds even where oracle/_ref is absent."""
import glob
import os

import numpy as np
import pytest
import torch

from curve_gaussian_b200.rasterizer import (GaussianRasterizationSettings, rasterize_backward_raw,
                                            rasterize_forward_raw)
from tests.test_gpu_raster_vs_reference import fetch, max_rel

pytestmark = pytest.mark.gpu
GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "raster_*.npz")))


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_matches_reference_golden(cuda_dev, path):
    z = np.load(path)
    dev = cuda_dev
    t = lambda k: torch.from_numpy(z[k]).to(dev)
    W, H = int(z["W"]), int(z["H"])
    rs = GaussianRasterizationSettings(H, W, float(z["tanx"]), float(z["tany"]),
                                       torch.full((3,), float(z["bg"]), device=dev), 1.0, t("viewmatrix"),
                                       t("projmatrix"), 0, t("campos"), False, False, False, True)
    R, color, radii, geom, bin_keep, img, invd, omap = rasterize_forward_raw(
        rs, t("means3D"), t("colors"), t("opacities"), t("scales"), t("rotations"), None, t("all_map"))
    scratch = rasterize_forward_raw.last_scratch
    P = radii.numel()
    assert R == int(z["R"])
    assert torch.equal(radii.cpu(), torch.from_numpy(z["radii"]))
    keys = fetch(0, P, R, W, H, geom, img, bin_keep, scratch, torch.int64, R)
    assert np.array_equal(keys.cpu().numpy(), z["keys"])
    pl = fetch(1, P, R, W, H, geom, img, bin_keep, scratch, torch.int32, R)
    assert np.array_equal(pl.cpu().numpy(), z["point_list"])
    nt = ((W + 15) // 16) * ((H + 15) // 16)
    rg = fetch(2, P, R, W, H, geom, img, bin_keep, scratch, torch.int32, 2 * nt)
    assert np.array_equal(rg.cpu().numpy().reshape(nt, 2), z["ranges"])
    nc = fetch(7, P, R, W, H, geom, img, bin_keep, scratch, torch.int32, W * H)
    assert np.array_equal(nc.cpu().numpy(), z["n_contrib"])
    assert np.array_equal(color.cpu().numpy(), z["color"])          # bit-identical pixels
    assert np.array_equal(invd.cpu().numpy(), z["invdepth"])
    assert np.array_equal(omap.cpu().numpy(), z["out_all_map"])
    bw = rasterize_backward_raw(rs, t("means3D"), radii, t("colors"), t("all_map"), t("opacities"), t("scales"),
                                t("rotations"), None, t("dL_dcolor"), t("dL_dinvdepth"), t("dL_dall_map_px"), geom, R,
                                bin_keep, img)
    names = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales",
             "dL_drotations", "dL_dall_map"]
    for n, g in zip(names, bw):
        if n == "dL_dsh":
            continue
        assert max_rel(g.cpu(), torch.from_numpy(z["ref_" + n])) <= 1e-5, n


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_block_contributor_lists(cuda_dev, path):
    """The per-4x4-block contributor lists the forward files for the ring backward (BinKeep::cand / cand_id, debug
    selectors 9-11): strictly ascending positions inside the tile's range, ids = the sorted point list at those
    positions, and the last entry of a block is the last contributor of its pixels (max n_contrib, which the golden
    test above pins to the reference)."""
    z = np.load(path)
    dev = cuda_dev
    t = lambda k: torch.from_numpy(z[k]).to(dev)
    W, H = int(z["W"]), int(z["H"])
    rs = GaussianRasterizationSettings(H, W, float(z["tanx"]), float(z["tany"]),
                                       torch.full((3,), float(z["bg"]), device=dev), 1.0, t("viewmatrix"),
                                       t("projmatrix"), 0, t("campos"), False, False, False, True)
    R, color, radii, geom, bin_keep, img, invd, omap = rasterize_forward_raw(
        rs, t("means3D"), t("colors"), t("opacities"), t("scales"), t("rotations"), None, t("all_map"))
    scratch = rasterize_forward_raw.last_scratch
    P = radii.numel()
    gx, gy = (W + 15) // 16, (H + 15) // 16
    nt = gx * gy
    cnt = fetch(9, P, R, W, H, geom, img, bin_keep, scratch, torch.int32, nt * 16).cpu().numpy().reshape(nt, 16)
    pos = fetch(10, P, R, W, H, geom, img, bin_keep, scratch, torch.int32, 16 * R).cpu().numpy()
    ids = fetch(11, P, R, W, H, geom, img, bin_keep, scratch, torch.int32, 16 * R).cpu().numpy()
    pl = z["point_list"]
    rg = z["ranges"]
    nc = np.zeros((gy * 16, gx * 16), np.int64)
    nc[:H, :W] = z["n_contrib"].reshape(H, W)
    checked = 0
    for tile in range(nt):
        x, y = int(rg[tile, 0]), int(rg[tile, 1])
        ty, tx = divmod(tile, gx)
        for b in range(16):
            n = int(cnt[tile, b])
            blk, half = b >> 1, b & 1
            px0, py0 = tx * 16 + (blk & 1) * 8 + half * 4, ty * 16 + (blk >> 1) * 4
            last = int(nc[py0:py0 + 4, px0:px0 + 4].max())
            if n == 0:
                assert last == 0, (tile, b)
                continue
            assert n <= y - x
            base = 16 * x + b * (y - x)
            p = pos[base:base + n].astype(np.int64)
            assert (np.diff(p) > 0).all() and p[0] >= 0 and p[-1] < y - x, (tile, b)
            assert np.array_equal(ids[base:base + n], pl[x + p]), (tile, b)
            assert p[-1] + 1 == last, (tile, b, p[-1], last)
            checked += 1
    assert checked > 0
