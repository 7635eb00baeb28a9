#!/usr/bin/env python
"""Where does the GPU idle inside a step? Runs a few bench steps under torch.profiler (CUPTI timestamps) and
prints the gaps between consecutive GPU activities (kernels, memsets, memcpys) that exceed a threshold."""
import json
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from curve_gaussian_b200 import synth  # noqa: E402
from curve_gaussian_b200.curve_model import GaussianCurveModel  # noqa: E402
from curve_gaussian_b200.loss import edge_ssim_loss  # noqa: E402
from curve_gaussian_b200.parallel import FlatGrad  # noqa: E402
from curve_gaussian_b200.renderer import render  # noqa: E402


class Pipe:
    debug = False
    antialiasing = False
    render_geo = True


def main():
    dev = torch.device("cuda", 0)
    B, n, W, H = 10000, 100, 1920, 1080
    cp, width, opl, isb = synth.random_curves(B, seed=0)
    model = GaussianCurveModel(0, n_gaussians=n, device=dev).create_from_curves(cp, width, opl, isb)
    bg = torch.zeros(3, device=dev)
    cams = [c.to(dev) for c in synth.random_cameras(4, W, H, seed=0)]
    gts = [torch.rand(1, H, W, device=dev) for _ in cams]
    fg = FlatGrad([model._curve_points, model._width, model._opacity, model._mask])

    def step(i):
        fg.flat.zero_()
        model.prepare_scaling_rot()
        image = render(cams[i % 4], model, Pipe(), bg)["render"]
        loss = edge_ssim_loss(image, gts[i % 4])
        loss.backward()

    for i in range(5):
        step(i)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for i in range(4):
            step(i)
        torch.cuda.synchronize()
    path = os.path.join(tempfile.gettempdir(), "trace.json")
    prof.export_chrome_trace(path)
    ev = json.load(open(path))["traceEvents"]
    gpu = [e for e in ev if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy")]
    gpu.sort(key=lambda e: e["ts"])
    t0, t1 = gpu[0]["ts"], gpu[-1]["ts"] + gpu[-1]["dur"]
    busy = sum(e["dur"] for e in gpu)
    print(f"GPU activities {len(gpu)}  span {t1 - t0:.0f} us  busy {busy:.0f} us  idle {t1 - t0 - busy:.0f} us  ({4} steps)")
    thr = float(os.environ.get("GAP_US", "8"))
    end = gpu[0]["ts"] + gpu[0]["dur"]
    prev = gpu[0]
    for e in gpu[1:]:
        gap = e["ts"] - end
        if gap > thr:
            print(f"  gap {gap:7.1f} us  after {prev['name'][:60]:60s} -> before {e['name'][:60]}")
        if e["ts"] + e["dur"] > end:
            end = e["ts"] + e["dur"]
            prev = e
    # host-side: time spent in cudaStreamSynchronize
    sync = [e for e in ev if e.get("ph") == "X" and "Synchronize" in e.get("name", "")]
    print("host sync calls", len(sync), "total us", sum(e["dur"] for e in sync))


if __name__ == "__main__":
    main()
