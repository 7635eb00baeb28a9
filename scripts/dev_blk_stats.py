"""Dev tool: contributor-list statistics of the 4x4 blocks (debug selector 9) on the C4 bench scene."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from curve_gaussian_b200 import _lib, synth
from curve_gaussian_b200 import rasterizer as rz
from curve_gaussian_b200.curve_model import GaussianCurveModel
from curve_gaussian_b200.renderer import render
import bench

dev = torch.device("cuda", 0)
B, n, W, H = 10000, 100, 1920, 1080
cp, width, opl, isb = synth.random_curves(B, seed=0)
model = GaussianCurveModel(0, n_gaussians=n, device=dev).create_from_curves(cp, width, opl, isb)
cam = synth.random_cameras(1, W, H, seed=0)[0].to(dev)
pipe = bench.Pipe()
hook = {}
orig = rz.rasterize_forward_raw
def spy(*a, **k):
    out = orig(*a, **k)
    hook["out"] = out
    return out
spy.last_scratch = None
rz.rasterize_forward_raw = spy
with torch.no_grad():
    render(cam, model, pipe, torch.zeros(3, device=dev))
R, color, radii, geom, bin_keep, img = hook["out"][:6]
lib = _lib.load()
nt = ((W + 15) // 16) * ((H + 15) // 16)
dst = torch.empty(nt * 16, dtype=torch.int32, device=dev)
_lib.check(lib.cg_raster_debug_fetch(9, radii.numel(), R, W, H, geom.data_ptr(), img.data_ptr(), bin_keep.data_ptr(),
                                     spy.last_scratch.data_ptr(), dst.data_ptr(), torch.cuda.current_stream().cuda_stream), "fetch")
c = dst.cpu().numpy().astype(np.int64)
print("R", R, "blocks", c.size, "non-empty", int((c > 0).sum()), "entries", int(c.sum()), "ideal warp-steps", int(c.sum()) // 2)
edges = [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 1 << 30]
for lo, hi in zip(edges[:-1], edges[1:]):
    m = (c >= lo) & (c < hi)
    print(f"[{lo},{hi}) blocks {int(m.sum())} entries {int(c[m].sum())} ({100.0 * c[m].sum() / c.sum():.1f}%)")
# epochs under the NB-starts-per-epoch rule, greedy in size order per ring: lower bound on steps = sum ceil-ish
for NB in (1, 2, 3, 4):
    s = np.sort(c[c > 0])[::-1]
    # blocks smaller than 16/NB waste slots when NB of them fill an epoch
    small = s[s < 16 // NB]
    waste = int((16 // NB * small.size - small.sum()))
    print("NB", NB, "slot waste bound (entries)", waste, f"{100.0 * waste / c.sum():.1f}%")
