import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_gpu_raster_vs_reference import *
dev = torch.device("cuda:0")
for case in ["cloud_small", "discs"]:
    cam, means, scales, rots, opac, colors, amap = make_case(case, dev)
    rs = settings_for(cam, dev, True, 0.0)
    H, W = rs.image_height, rs.image_width; P = means.shape[0]
    z1 = torch.zeros(1,H,W,device=dev); z4 = torch.zeros(4,H,W,device=dev)
    (R_ref, color_ref, radii_ref, geomB, binB, imgB, invd_ref, omap_ref), bw_ref = run_reference(rs, means, colors, opac, scales, rots, amap, (z1, z1, z4))
    R, color, radii, geom, bin_keep, img, invd, omap = rasterize_forward_raw(rs, means, colors, opac, scales, rots, None, amap)
    scratch = rasterize_forward_raw.last_scratch
    dec = decode_ref_buffers(geomB, binB, imgB, P, R_ref, W * H)
    vis = radii_ref > 0
    co = fetch(6, P, R, W, H, geom, img, bin_keep, scratch, torch.float32, 4 * P).view(P, 4)
    rc = dec["conic_opacity"].view(P, 4)
    for c in range(4):
        d = (co[vis][:, c].view(torch.int32) - rc[vis][:, c].view(torch.int32))
        print(case, "comp", c, "mismatch", (d != 0).sum().item(), "of", vis.sum().item(), "max ulp", d.abs().max().item())
    print(case, "color bit-equal:", torch.equal(color, color_ref), "max abs", (color-color_ref).abs().max().item())
