"""Dev tool: where the HOST time of a small-scene step goes (cProfile, default-start shape)."""
import cProfile, pstats, os, sys, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from curve_gaussian_b200 import synth
from curve_gaussian_b200.curve_model import GaussianCurveModel
from curve_gaussian_b200.loss import edge_ssim_loss
from curve_gaussian_b200.renderer import render
dev = torch.device("cuda:0")
class Pipe:
    debug = False; antialiasing = False; render_geo = True
B, n, W, H = 3375, 12, 800, 800
cp, width, opl, isb = synth.random_curves(B, seed=0)
model = GaussianCurveModel(0, n_gaussians=n, device=dev).create_from_curves(cp, width, opl, isb)
bg = torch.zeros(3, device=dev)
cams = [c.to(dev) for c in synth.random_cameras(4, W, H, seed=0)]
gts = [torch.rand(1, H, W, device=dev) for _ in cams]
pipe = Pipe()
def step(i):
    for p in model.parameters():
        p.grad = None
    model.prepare_scaling_rot()
    image = render(cams[i % 4], model, pipe, bg)["render"]
    loss = edge_ssim_loss(image, gts[i % 4])
    loss.backward()
for i in range(20): step(i)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for i in range(200): step(i)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45); print(s.getvalue()[:9000])
