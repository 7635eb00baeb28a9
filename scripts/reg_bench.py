"""Dev tool: fused curve regularisers vs the reference's torch formulation at the C4 scene size (B=10k, n=100)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from curve_gaussian_b200 import synth
from curve_gaussian_b200.curve_model import GaussianCurveModel
from curve_gaussian_b200.regularizers import curve_smoothness, endpoint_connectivity
from oracle import torch_ref
dev = torch.device("cuda:0")
B, n = 10000, 100
cp, width, opl, isb = synth.random_curves(B, seed=0)
model = GaussianCurveModel(0, n_gaussians=n, device=dev).create_from_curves(cp, width, opl, isb)
def timeit(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
q = model._rotation.detach().clone().requires_grad_(True)
x = model._curve_points.detach().clone().requires_grad_(True)
def ours_s(): q.grad = None; curve_smoothness(q, n).backward()
def ref_s(): q.grad = None; torch_ref.curve_smoothness(q, n).backward()
def ours_c(): x.grad = None; endpoint_connectivity(x).backward()
def ref_c(): x.grad = None; torch_ref.endpoint_connectivity(x).backward()
print(f"curve smoothness fwd+bwd  : fused {timeit(ours_s):.3f} ms   torch formulation {timeit(ref_s):.3f} ms")
torch.cuda.reset_peak_memory_stats(); m0 = torch.cuda.memory_allocated()
t = timeit(ours_c); pk = torch.cuda.max_memory_allocated() - m0
print(f"endpoint connectivity     : fused {t:.3f} ms (peak extra memory {pk/1e6:.1f} MB)")
torch.cuda.reset_peak_memory_stats(); m0 = torch.cuda.memory_allocated()
t = timeit(ref_c, it=3); pk = torch.cuda.max_memory_allocated() - m0
print(f"                            torch formulation {t:.3f} ms (peak extra memory {pk/1e6:.1f} MB)")
print("values:", curve_smoothness(q, n).item(), torch_ref.curve_smoothness(q, n).item(), endpoint_connectivity(x).item(), torch_ref.endpoint_connectivity(x).item())
