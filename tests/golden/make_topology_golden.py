"""Generates tests/golden/topology_*.npz by running the REFERENCE's own curve-set surgery
(scene/gaussian_curve_model.py: de_casteljau_split/trim, densify_and_prune, curve_split_curvature, only_prune,
mask_trim_split, reset_opacity, prune_curves + the Adam-state edits of scene/gaussian_model.py) on CPU in the
build container. The reference hard-codes device="cuda" in these methods; the factories it calls are wrapped so
that "cuda" means the CPU here - the arithmetic is untouched.
Run:  python tests/golden/make_topology_golden.py     (needs /root/reference; the .npz files are committed)
"""
import os
import sys
import types

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
from curve_gaussian_b200 import synth  # noqa: E402
from make_sampling_golden import import_reference_model  # noqa: E402


def cuda_means_cpu():
    """Wrap the tensor factories the reference calls with device='cuda'."""
    def wrap(fn):
        def inner(*a, **k):
            if str(k.get("device", "")).startswith("cuda"):
                k["device"] = "cpu"
            return fn(*a, **k)
        return inner
    for name in ("zeros", "ones", "linspace", "tensor", "eye", "empty", "zeros_like", "ones_like"):
        setattr(torch, name, wrap(getattr(torch, name)))
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.empty_cache = lambda: None


class Args:
    feature_lr = 0.0025
    opacity_lr = 0.05
    scaling_lr = 0.005
    mask_lr = 0.01
    lr_curve_points_init = 0.00016
    lr_curve_points_final = 0.0000016
    position_lr_delay_mult = 0.01
    position_lr_max_steps = 30000
    exposure_lr_init = 0.01
    exposure_lr_final = 0.001
    exposure_lr_delay_steps = 0
    exposure_lr_delay_mult = 0.0
    iterations = 30000


def build(GaussianCurveModel, B, n, seed, line_fraction):
    cp, width, opl, isb = synth.random_curves(B, seed=seed, line_fraction=line_fraction)
    g = torch.Generator().manual_seed(seed + 1)
    m = GaussianCurveModel.__new__(GaussianCurveModel)      # __init__ hard-codes device='cuda'
    m.n_gaussians = n
    m.max_sh_degree = 0
    m.active_sh_degree = 0
    m.optimizer_type = "default"
    m.sample_t = torch.linspace(0.5 / n, 1 - 0.5 / n, n)[:, None, None]
    m.opacity_activation = torch.sigmoid
    m.inverse_opacity_activation = lambda x: torch.log(x / (1 - x))
    m.scaling_activation = torch.exp
    m.scaling_inverse_activation = torch.log
    m.rotation_activation = torch.nn.functional.normalize
    m._curve_points = nn.Parameter(cp.clone().requires_grad_(True))
    m._width = nn.Parameter((width + 0.3 * torch.randn(B, 1, generator=g)).requires_grad_(True))
    m._opacity = nn.Parameter((opl + 1.5 * torch.randn(B, 1, generator=g)).requires_grad_(True))
    m._mask = nn.Parameter((2.0 + 3.0 * torch.randn(B, n, 1, generator=g)).requires_grad_(True))
    m._features_dc = nn.Parameter(torch.zeros(B, n, 1, 1).requires_grad_(True))
    m._features_rest = nn.Parameter(torch.zeros(B, n, 0, 1).requires_grad_(True))
    m._exposure = nn.Parameter(torch.eye(3, 4)[None].clone().requires_grad_(True))
    m.pretrained_exposures = None
    m.is_bezier = isb.clone()
    m.max_radii2D = torch.zeros(B * n)
    m.spatial_lr_scale = 1.0
    m.prepare_scaling_rot()
    m.training_setup(Args())
    # one Adam step with random gradients so that every group has moments to carry through the surgery
    for p in (m._curve_points, m._width, m._opacity, m._mask, m._features_dc):
        p.grad = 0.1 * torch.randn(p.shape, generator=g)
    m._features_rest.grad = torch.zeros_like(m._features_rest)
    m.optimizer.step()
    m.xyz_gradient_accum = torch.rand(B * n, 1, generator=g) * 1e-3
    m.denom = torch.randint(0, 4, (B * n, 1), generator=g).float()     # some zeros -> NaN -> 0 path
    m.prepare_scaling_rot()
    return m, g


def state(m, prefix):
    out = {prefix + "curve_points": m._curve_points, prefix + "width": m._width, prefix + "opacity": m._opacity,
           prefix + "mask": m._mask, prefix + "is_bezier": m.is_bezier, prefix + "accum": m.xyz_gradient_accum,
           prefix + "denom": m.denom, prefix + "max_radii2D": m.max_radii2D}
    for group in m.optimizer.param_groups:
        st = m.optimizer.state.get(group["params"][0])
        if st is not None and group["name"] in ("curve_points", "width", "opacity", "mask"):
            out[prefix + "exp_avg_" + group["name"]] = st["exp_avg"]
            out[prefix + "exp_avg_sq_" + group["name"]] = st["exp_avg_sq"]
    return {k: v.detach().cpu().numpy().copy() for k, v in out.items()}   # copy: some methods edit in place


def main():
    GaussianCurveModel = import_reference_model()
    cuda_means_cpu()
    # 1. pure split / trim
    cp, _, _, isb = synth.random_curves(40, seed=3, line_fraction=0.3)
    g = torch.Generator().manual_seed(9)
    t = torch.rand(40, 1, generator=g) * 0.9 + 0.05
    t2 = torch.rand(40, 1, generator=g) * 0.5 + 0.4
    m = GaussianCurveModel.__new__(GaussianCurveModel)
    m.is_bezier = isb
    left, right = m.de_casteljau_split(cp, t, isb)
    trimmed = m.de_casteljau_trim(cp, t * 0.5, t2, isb)
    m.is_bezier = torch.ones(40, dtype=torch.bool)
    left_b, right_b = m.de_casteljau_split(cp, t, m.is_bezier)
    np.savez_compressed(os.path.join(HERE, "topology_split.npz"), cp=cp.numpy(), is_bezier=isb.numpy(), t=t.numpy(),
                        t2=t2.numpy(), left=left.numpy(), right=right.numpy(), trimmed=trimmed.numpy(),
                        left_b=left_b.numpy(), right_b=right_b.numpy())

    # 2. the surgery calls, each on a fresh model
    cases = {
        "densify": lambda m, g: m.densify_and_prune(2.5e-4, 0.35, 1.0, 20, torch.ones(m._xyz.shape[0])),
        "curvature": lambda m, g: m.curve_split_curvature(threshold_angle=6, threshold_radian_skip=10),
        "only_prune": lambda m, g: m.only_prune(0.4, 0.6),
        "mask_trim": lambda m, g: m.mask_trim_split(0.7),
        "reset_opacity": lambda m, g: m.reset_opacity(),
        "fix_opacity": lambda m, g: m.fix_opacity(),
    }
    # 3. fit_curve_to_line: nearly straight Beziers become lines (decision + unchanged points + reset moments)
    m, g = build(GaussianCurveModel, 60, 12, seed=33, line_fraction=0.2)
    with torch.no_grad():
        cp = m._curve_points
        chord = cp[:, 3] - cp[:, 0]
        flat = torch.stack([cp[:, 0], cp[:, 0] + chord / 3, cp[:, 0] + 2 * chord / 3, cp[:, 3]], dim=1)
        bend = torch.linspace(0, 1, 60)[:, None, None] ** 2          # from exactly straight to the original shape
        cp.copy_(flat + bend * (cp - flat))
    before = state(m, "in_")
    with torch.no_grad():                       # train.py calls it inside its no_grad block (:166, :213-215)
        m.fit_curve_to_line(0.002, 0.004)
    after = state(m, "out_")
    np.savez_compressed(os.path.join(HERE, "topology_fit_line.npz"), n=12, **before, **after)
    print("fit_line", int(before["in_is_bezier"].sum()), "->", int(after["out_is_bezier"].sum()), "beziers")

    # 4. merge_curves with the RANSAC consensus step replaced by "every point is an inlier" (skimage is not in this
    #    image and the reference does not seed it): everything else - pairing, ordering, scipy's curve_fit, line
    #    components - is the reference's own code
    import scene.gaussian_curve_model as ref_mod
    ref_mod.ransac = lambda pts, model, min_samples, residual_threshold, max_trials: (None, np.ones(len(pts), dtype=bool))
    m, g = build(GaussianCurveModel, 24, 12, seed=41, line_fraction=0.0)
    with torch.no_grad():
        base, _, _, _ = synth.random_curves(14, seed=43)
        halves = []
        for c in base[:8]:                                   # 8 smooth curves cut in two -> 16 mergeable halves
            l, r = m.de_casteljau_split(c[None], torch.tensor([[0.5]]), torch.ones(1, dtype=torch.bool))
            halves += [l[0], r[0]]
        segs = []
        for c in base[8:14]:                                 # 6 chords cut in three collinear segments
            a, b = c[0], c[3]
            for k in range(3):
                p, q = a + (b - a) * k / 3, a + (b - a) * (k + 1) / 3
                segs.append(torch.stack([p, p + (q - p) / 3, p + 2 * (q - p) / 3, q]))
        extra = m._curve_points[:24 - 16].detach().clone()   # unrelated curves
        allc = torch.stack(halves + segs + list(extra))
        perm = torch.randperm(allc.shape[0], generator=g)
        isb = torch.tensor([True] * 16 + [False] * 18 + [True] * extra.shape[0])[perm]
        allc = allc[perm]
    B2 = allc.shape[0]
    m2, g = build(GaussianCurveModel, B2, 12, seed=45, line_fraction=0.0)
    with torch.no_grad():
        m2._curve_points.copy_(allc)
    m2.is_bezier = isb.clone()
    m2.prepare_scaling_rot()
    before = state(m2, "in_")
    with torch.no_grad():
        m2.merge_curves(0.02, 0.97)
    after = state(m2, "out_")
    np.savez_compressed(os.path.join(HERE, "topology_merge.npz"), n=12, **before, **after)
    print("merge", before["in_curve_points"].shape[0], "->", after["out_curve_points"].shape[0], "curves,",
          int(after["out_is_bezier"].sum()), "beziers")

    for name, fn in cases.items():
        m, g = build(GaussianCurveModel, 60, 12, seed=21, line_fraction=0.25)
        before = state(m, "in_")
        with torch.no_grad() if name in ("mask_trim",) else torch.enable_grad():
            fn(m, g)
        after = state(m, "out_")
        after["out_xyz"] = m._xyz.detach().numpy()
        np.savez_compressed(os.path.join(HERE, f"topology_{name}.npz"), n=12, **before, **after)
        print(name, before["in_curve_points"].shape[0], "->", after["out_curve_points"].shape[0], "curves")


if __name__ == "__main__":
    main()
