"""Dev tool: full-step time (sample -> render -> fused loss -> backward) at the smaller BASELINE.json shapes,
where launch/host overhead rather than kernel time decides the step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from curve_gaussian_b200 import synth, _lib
from curve_gaussian_b200.curve_model import GaussianCurveModel
from curve_gaussian_b200.loss import edge_ssim_loss
from curve_gaussian_b200.renderer import render

dev = torch.device("cuda:0")
class Pipe:
    debug = False; antialiasing = False; render_geo = True

def run(name, B, n, W, H, steps=50):
    cp, width, opl, isb = synth.random_curves(B, seed=0)
    model = GaussianCurveModel(0, n_gaussians=n, device=dev).create_from_curves(cp, width, opl, isb)
    bg = torch.zeros(3, device=dev)
    cams = [c.to(dev) for c in synth.random_cameras(4, W, H, seed=0)]
    gts = [torch.rand(1, H, W, device=dev) for _ in cams]
    def step(i):
        for p in model.parameters():
            p.grad = None
        model.prepare_scaling_rot()
        image = render(cams[i % 4], model, Pipe(), bg)["render"]
        loss = edge_ssim_loss(image, gts[i % 4])
        loss.backward()
    for i in range(5): step(i)
    torch.cuda.synchronize()
    lib = _lib.load(); lib.cg_profile_reset(); 
    t0 = time.perf_counter()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps): step(i)
    e1.record(); torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / steps * 1e3
    gpu = e0.elapsed_time(e1) / steps
    lib.cg_profile_enable(1)
    for i in range(10): step(i)
    torch.cuda.synchronize(); lib.cg_profile_enable(0)
    st = _lib.profile_read(); ksum = sum(v[0] for v in st.values()) / 10
    print(f"{name:28s} P={B*n:8d} {W}x{H}: step {gpu:.3f} ms (host wall {wall:.3f} ms), libcurvegs kernels {ksum:.3f} ms", flush=True)

run("default start 3375x12 @800^2", 3375, 12, 800, 800)
run("C2 shape 417x12 @800^2", 417, 12, 800, 800)
run("C3 shape 8334x12 @1200x680", 8334, 12, 1200, 680)
run("C4 10000x100 @1920x1080", 10000, 100, 1920, 1080, steps=20)
