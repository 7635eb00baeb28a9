"""Dev tool: the C2-shaped step (417 x 12 curve-Gaussians, 800x800) eager, for an ncu launch list of its kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from curve_gaussian_b200 import synth
from curve_gaussian_b200.curve_model import GaussianCurveModel
from curve_gaussian_b200.loss import edge_ssim_loss
from curve_gaussian_b200.parallel import FlatGrad
from curve_gaussian_b200.renderer import render

dev = torch.device("cuda", 0)
B, n, W, H = 417, 12, 800, 800
cp, width, opl, isb = synth.random_curves(B, seed=0)
model = GaussianCurveModel(0, n_gaussians=n, device=dev).create_from_curves(cp, width, opl, isb)
fg = FlatGrad([model._curve_points, model._width, model._opacity, model._mask], direct=True)
bg = torch.zeros(3, device=dev)
cams = [c.to(dev) for c in synth.random_cameras(4, W, H, seed=0)]
gts = [torch.rand(1, H, W, device=dev) for _ in cams]
pipe = bench.Pipe()
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    fg.zero()
    model.prepare_scaling_rot()
    loss = edge_ssim_loss(render(cams[i % 4], model, pipe, bg)["render_raw"], gts[i % 4], clamp=True)
    loss.backward()
torch.cuda.synchronize()
print("ok", float(loss))
