"""Generates tests/golden/raster_*.npz on a B200 by running the UNMODIFIED reference
CUDA rasterizer (oracle/_ref/diff_cur_rasterization_C.so, built from /root/reference by
oracle/build_ref.sh). Run on the GPU box:
    python tests/golden/make_raster_golden.py gpurun_out/golden
then copy gpurun_out/golden/*.npz into tests/golden/. The committed .npz files are what
tests/test_oracle_raster.py (CPU) and tests/test_gpu_raster_golden.py (GPU) check against.
"""
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from curve_gaussian_b200 import synth  # noqa: E402
from oracle import torch_ref  # noqa: E402
from tests import refload  # noqa: E402
from tests.test_gpu_raster_vs_reference import decode_ref_buffers  # noqa: E402


def case_inputs(name):
    if name == "c1_bezier32":   # BASELINE.json configs[0]
        W = H = 128
        cp, width, opl, isb = synth.random_curves(1, seed=0)
        cp = (cp - 0.5) * 2.0 + 0.5   # one long curve across the view
        n = 32
        xyz, q, sc = torch_ref.sample_curves(cp, width + 1.2, isb, n)
        cam = synth.random_cameras(1, W, H, seed=3)[0]
        m3, op, scl, rot, col, amap = torch_ref.raster_inputs(xyz, q, sc, opl, n, torch.ones(1, n, 1),
                                                              cam.camera_center, cam.world_view_transform)
        bgv = 0.0
    elif name == "cloud400":
        W, H = 100, 75
        m3, scl, rot, op, col, amap = synth.random_gaussians(400, seed=21, scale_lo=0.01, scale_hi=0.06)
        cam = synth.random_cameras(1, W, H, seed=5)[0]
        bgv = 0.3
    elif name == "discs1500":
        W, H = 160, 120
        m3, scl, rot, op, col, amap = synth.random_gaussians(1500, seed=22)
        P = 1500
        scl = torch.stack([torch.full((P,), 2e-3), torch.full((P,), 1.5e-2), torch.full((P,), 1.5e-2)], 1)
        rot = rot * (0.7 + 0.6 * torch.rand(P, 1, generator=torch.Generator().manual_seed(6)))
        op = torch.full((P, 1), 0.6)
        col = torch.ones(P, 1)
        cam = synth.random_cameras(1, W, H, seed=6)[0]
        bgv = 0.0
    else:
        raise KeyError(name)
    return W, H, cam, bgv, [t.detach().contiguous() for t in (m3, scl, rot, op, col, amap)]


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    dev = torch.device("cuda:0")
    ref = refload.ref_rasterizer()
    assert ref is not None, "build oracle/_ref first"
    empty = torch.Tensor([])
    for name in ("c1_bezier32", "cloud400", "discs1500"):
        W, H, cam, bgv, (m3, scl, rot, op, col, amap) = case_inputs(name)
        tanx, tany = math.tan(cam.FoVx * 0.5), math.tan(cam.FoVy * 0.5)
        g = torch.Generator().manual_seed(11)
        dL_c = torch.randn(1, H, W, generator=g)
        dL_d = torch.randn(1, H, W, generator=g) * 0.1
        dL_m = torch.randn(4, H, W, generator=g) * 0.1
        d = lambda t: t.to(dev)
        bg = torch.full((3,), bgv, device=dev)
        out = ref.rasterize_gaussians(bg, d(m3), d(col), d(op), d(scl), d(rot), 1.0, empty, d(amap),
                                      d(cam.world_view_transform), d(cam.full_proj_transform), tanx, tany, H, W,
                                      empty, 0, d(cam.camera_center), False, False, True, False)
        R, color, radii, geomB, binB, imgB, invd, omap = out
        bw = ref.rasterize_gaussians_backward(bg, omap, d(m3), radii, d(col), d(amap), d(op), d(scl), d(rot), 1.0,
                                              empty, d(cam.world_view_transform), d(cam.full_proj_transform), tanx,
                                              tany, d(dL_c), d(dL_d), d(dL_m), empty, 0, d(cam.camera_center), geomB,
                                              R, binB, imgB, False, True, False)
        torch.cuda.synchronize()
        P = m3.shape[0]
        dec = decode_ref_buffers(geomB, binB, imgB, P, R, W * H)
        ntiles = ((W + 15) // 16) * ((H + 15) // 16)
        c = lambda t: t.detach().cpu().numpy()
        names = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales",
                 "dL_drotations", "dL_dall_map"]
        save = dict(W=np.int64(W), H=np.int64(H), tanx=np.float64(tanx), tany=np.float64(tany), bg=np.float32(bgv),
                    means3D=c(m3), scales=c(scl), rotations=c(rot), opacities=c(op), colors=c(col), all_map=c(amap),
                    viewmatrix=c(cam.world_view_transform), projmatrix=c(cam.full_proj_transform),
                    campos=c(cam.camera_center), dL_dcolor=c(dL_c), dL_dinvdepth=c(dL_d), dL_dall_map_px=c(dL_m),
                    R=np.int64(R), radii=c(radii), color=c(color), invdepth=c(invd), out_all_map=c(omap),
                    depths=c(dec["depths"]), means2D=c(dec["means2D"]).reshape(P, 2),
                    conic_opacity=c(dec["conic_opacity"]).reshape(P, 4), tiles_touched=c(dec["tiles_touched"]),
                    n_contrib=c(dec["n_contrib"]), final_T=c(dec["accum_alpha"]),
                    ranges=c(dec["ranges"])[:2 * ntiles].reshape(ntiles, 2),
                    keys=c(dec["keys"]) if R > 0 else np.zeros(0, np.int64),
                    point_list=c(dec["point_list"]) if R > 0 else np.zeros(0, np.int32))
        for n_, t in zip(names, bw):
            if n_ != "dL_dsh":
                save["ref_" + n_] = c(t)
        np.savez_compressed(os.path.join(outdir, f"raster_{name}.npz"), **save)
        print(name, "P", P, "R", R, "saved")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
