"""Dev tool: full-step time (sample -> render -> fused loss -> backward) at the BASELINE.json shapes, three ways:
eager with the exact binning (the reference's host round trip for R), eager with sync-free capacity binning, and
the whole step replayed from a CUDA graph. Prints one JSON line per shape."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from curve_gaussian_b200 import synth, _lib
from curve_gaussian_b200 import rasterizer as rz
from curve_gaussian_b200.curve_model import GaussianCurveModel
from curve_gaussian_b200.graph import GraphedStep, StaticCamera
from curve_gaussian_b200.loss import edge_ssim_loss
from curve_gaussian_b200.parallel import FlatGrad
from curve_gaussian_b200.renderer import render

dev = torch.device("cuda:0")
class Pipe:
    debug = False; antialiasing = False; render_geo = True

def timed(fn, steps, warm=5):
    for i in range(warm): fn(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps): fn(warm + i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, (time.perf_counter() - t0) / steps * 1e3

def run(name, B, n, W, H, steps=100, nviews=8):
    cp, width, opl, isb = synth.random_curves(B, seed=0)
    model = GaussianCurveModel(0, n_gaussians=n, device=dev).create_from_curves(cp, width, opl, isb)
    fg = FlatGrad([model._curve_points, model._width, model._opacity, model._mask])
    bg = torch.zeros(3, device=dev)
    cams = [c.to(dev) for c in synth.random_cameras(nviews, W, H, seed=0)]
    gts = [torch.rand(1, H, W, device=dev) for _ in cams]
    scam = StaticCamera(cams[0]); gt_static = torch.empty_like(gts[0])

    def body(cam, gt):
        fg.zero()
        model.prepare_scaling_rot()
        loss = edge_ssim_loss(render(cam, model, Pipe(), bg)["render_raw"], gt, clamp=True)
        loss.backward()
        return loss
    eager = lambda i: body(cams[i % nviews], gts[i % nviews])
    ms_exact, wall_exact = timed(eager, steps)
    lib = _lib.load(); lib.cg_profile_reset(); lib.cg_profile_enable(1)
    for i in range(10): eager(i)
    torch.cuda.synchronize(); lib.cg_profile_enable(0)
    ksum = sum(v[0] for v in _lib.profile_read().values()) / 10
    pol = rz.CapacityBinning()
    with rz.capacity_binning(pol):
        def eager_cap(i):
            r = eager(i); pol.poll(); return r
        ms_cap, wall_cap = timed(eager_cap, steps)
        torch.cuda.synchronize(); pol.check()
    gs = GraphedStep(lambda: body(scam, gt_static), policy=pol,
                     calibrate=[(lambda c=c: scam.load(c)) for c in cams]).capture()
    def graphed(i):
        scam.load(cams[i % nviews]); gt_static.copy_(gts[i % nviews], non_blocking=True)
        return gs.replay()
    ms_graph, wall_graph = timed(graphed, steps)
    ok = gs.verify()
    print(json.dumps({"shape": name, "P": B * n, "image": f"{W}x{H}", "kernels_ms": round(ksum, 4),
                      "eager_exact_ms": round(ms_exact, 4), "eager_capacity_ms": round(ms_cap, 4),
                      "graph_ms": round(ms_graph, 4), "graph_views_per_s": round(1e3 / ms_graph, 1),
                      "host_wall_ms": {"exact": round(wall_exact, 4), "capacity": round(wall_cap, 4), "graph": round(wall_graph, 4)},
                      "capacity": {str(k): v for k, v in pol.caps.items()}, "max_R": {str(k): v for k, v in pol.max_seen.items()}, "overflows": pol.overflows, "verified": ok},
                     default=str), flush=True)

if __name__ == "__main__":
    which = sys.argv[1:] or ["start", "c2", "c3", "c4"]
    if "start" in which: run("default start 3375x12 @800^2", 3375, 12, 800, 800)
    if "c2" in which: run("C2 shape 417x12 @800^2", 417, 12, 800, 800)
    if "c3" in which: run("C3 shape 8334x12 @1200x680", 8334, 12, 1200, 680)
    if "c4" in which: run("C4 10000x100 @1920x1080", 10000, 100, 1920, 1080, steps=30)
