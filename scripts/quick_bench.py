"""Dev tool: time our rasterizer fwd+bwd vs the recompiled reference CUDA on a C4-shaped scene."""
import math, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from curve_gaussian_b200 import synth
from curve_gaussian_b200.rasterizer import GaussianRasterizationSettings, rasterize_forward_raw, rasterize_backward_raw
from oracle import torch_ref
from tests import refload

dev = torch.device("cuda:0")
B = int(os.environ.get("B", 10000)); n = int(os.environ.get("N", 100))
W = int(os.environ.get("W", 1920)); H = int(os.environ.get("H", 1080))
cp, width, opl, isb = synth.random_curves(B, seed=0)
cp, width, opl, isb = cp.to(dev), width.to(dev), opl.to(dev), isb.to(dev)
cams = [c.to(dev) for c in synth.random_cameras(4, W, H, seed=0)]
with torch.no_grad():
    xyz, q, sc = torch_ref.sample_curves(cp, width, isb, n)
P = xyz.shape[0]
mask = torch.ones(B, n, 1, device=dev)

def timeit(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

for ci, cam in enumerate(cams[:2]):
    with torch.no_grad():
        m3, op, scl, rot, col, amap = torch_ref.raster_inputs(xyz, q, sc, opl, n, mask, cam.camera_center, cam.world_view_transform)
    rs = GaussianRasterizationSettings(H, W, math.tan(cam.FoVx/2), math.tan(cam.FoVy/2), torch.zeros(3, device=dev), 1.0,
        cam.world_view_transform, cam.full_proj_transform, 0, cam.camera_center, False, False, False, True)
    g_color = torch.randn(1, H, W, device=dev)
    out = rasterize_forward_raw(rs, m3, col, op, scl, rot, None, amap)
    R = out[0]
    print(f"cam {ci}: P={P} R={R} R/P={R/P:.2f} visible={(out[2]>0).sum().item()} max_radius={out[2].max().item()}")
    t_f = timeit(lambda: rasterize_forward_raw(rs, m3, col, op, scl, rot, None, amap))
    R, color, radii, geom, bk, img, invd, omap = out
    t_b = timeit(lambda: rasterize_backward_raw(rs, m3, radii, col, amap, op, scl, rot, None, g_color, None, None, geom, R, bk, img))
    print(f"  ours: fwd {t_f:.3f} ms  bwd(color-only) {t_b:.3f} ms")
    zeros1 = torch.zeros(1, H, W, device=dev); zeros4 = torch.zeros(4, H, W, device=dev)
    t_b2 = timeit(lambda: rasterize_backward_raw(rs, m3, radii, col, amap, op, scl, rot, None, g_color, zeros1, zeros4, geom, R, bk, img))
    print(f"  ours: bwd(all channels) {t_b2:.3f} ms")
    ref = refload.ref_rasterizer()
    if ref is not None:
        empty = torch.Tensor([])
        def rf():
            return ref.rasterize_gaussians(rs.bg, m3, col, op, scl, rot, 1.0, empty, amap, rs.viewmatrix, rs.projmatrix,
                rs.tanfovx, rs.tanfovy, H, W, empty, 0, rs.campos, False, False, True, False)
        o = rf()
        Rr, colr, radr, gB, bB, iB, invr, omr = o
        def rb():
            return ref.rasterize_gaussians_backward(rs.bg, omr, m3, radr, col, amap, op, scl, rot, 1.0, empty, rs.viewmatrix,
                rs.projmatrix, rs.tanfovx, rs.tanfovy, g_color, zeros1, zeros4, empty, 0, rs.campos, gB, Rr, bB, iB, False, True, False)
        print(f"  ref : R={Rr} fwd {timeit(rf):.3f} ms  bwd {timeit(rb):.3f} ms")
        print("  color max abs diff", (colr - color).abs().max().item(), "radii equal", torch.equal(radr, radii))
