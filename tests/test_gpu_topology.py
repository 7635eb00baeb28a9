"""GPU: the curve-set surgery and optimizer bookkeeping run on the device - with the CUDA sampling
op re-sampling the curves after every edit - against the same golden vectors the CPU suite uses, which were produced
by the reference's OWN methods (tests/golden/make_topology_golden.py). Same decisions (curve counts, kinds, which
curves are split / pruned) and the same values: the edits are elementwise, so they are compared bit for bit except
where the CPU test already documents a tolerance."""
import numpy as np
import pytest
import torch

from curve_gaussian_b200 import topology
from curve_gaussian_b200.curve_model import GaussianCurveModel
from tests.test_topology_io import Args, GOLD, load

pytestmark = pytest.mark.gpu


def model_from(d, dev):
    t = lambda k: torch.from_numpy(d["in_" + k]).to(dev)
    n = int(d["n"])
    m = GaussianCurveModel(0, n_gaussians=n, device=dev).create_from_curves(t("curve_points"), t("width"), t("opacity"),
                                                                            t("is_bezier"), t("mask"))
    m.training_setup(Args())
    for group in m.optimizer.param_groups:
        p = group["params"][0]
        st = {"step": torch.tensor(1.0), "exp_avg": torch.zeros_like(p), "exp_avg_sq": torch.zeros_like(p)}
        if "in_exp_avg_" + group["name"] in d:
            st["exp_avg"] = t("exp_avg_" + group["name"]).clone()
            st["exp_avg_sq"] = t("exp_avg_sq_" + group["name"]).clone()
        m.optimizer.state[p] = st
    m.xyz_gradient_accum, m.denom = t("accum").clone(), t("denom").clone()
    m.max_radii2D = t("max_radii2D").clone()
    return m


def check(m, d, mask_tol=0.0, opacity_tol=0.0):
    t = lambda k: torch.from_numpy(d["out_" + k])
    c = lambda x: x.detach().cpu()
    assert m._curve_points.shape == t("curve_points").shape
    assert m._curve_points.is_cuda and m._xyz.is_cuda
    assert torch.equal(c(m.is_bezier), t("is_bezier"))
    for name, attr in (("curve_points", "_curve_points"), ("width", "_width"), ("opacity", "_opacity")):
        if name == "opacity" and opacity_tol:
            assert (c(getattr(m, attr)) - t(name)).abs().max() <= opacity_tol, name
        else:
            assert torch.equal(c(getattr(m, attr)), t(name)), name
    if mask_tol:
        assert (c(m._mask) - t("mask")).abs().max() <= mask_tol
    else:
        assert torch.equal(c(m._mask), t("mask"))
    assert torch.equal(c(m.xyz_gradient_accum), t("accum")) and torch.equal(c(m.denom), t("denom"))
    assert torch.equal(c(m.max_radii2D), t("max_radii2D"))
    for group in m.optimizer.param_groups:
        key = "out_exp_avg_" + group["name"]
        if key in d:
            st = m.optimizer.state[group["params"][0]]
            assert torch.equal(c(st["exp_avg"]), torch.from_numpy(d[key])), key
            assert torch.equal(c(st["exp_avg_sq"]), torch.from_numpy(d["out_exp_avg_sq_" + group["name"]])), key
    # re-sampled by the CUDA op after the surgery
    assert m._xyz.shape[0] == m._curve_points.shape[0] * m.n_gaussians
    assert torch.isfinite(m._xyz).all() and torch.isfinite(m._rotation).all() and torch.isfinite(m._scaling).all()


def test_de_casteljau_split_and_trim_on_device(cuda_dev):
    d = np.load(GOLD + "/topology_split.npz")
    cp, isb, t, t2 = (torch.from_numpy(d[k]).to(cuda_dev) for k in ("cp", "is_bezier", "t", "t2"))
    left, right = topology.de_casteljau_split(cp, t, isb)
    assert torch.equal(left.cpu(), torch.from_numpy(d["left"])) and torch.equal(right.cpu(), torch.from_numpy(d["right"]))
    assert torch.equal(topology.de_casteljau_trim(cp, t * 0.5, t2, isb).cpu(), torch.from_numpy(d["trimmed"]))


@pytest.mark.parametrize("name", ["densify", "curvature", "only_prune", "reset_opacity", "fix_opacity"])
def test_surgery_on_device_matches_reference(cuda_dev, name):
    d = load(name)
    m = model_from(d, cuda_dev)
    n_before = m._curve_points.shape[0]
    {"densify": lambda: m.densify_and_prune(2.5e-4, 0.35, 1.0, 20, torch.ones(m._xyz.shape[0], device=cuda_dev)),
     "curvature": lambda: m.curve_split_curvature(threshold_angle=6, threshold_radian_skip=10),
     "only_prune": lambda: m.only_prune(0.4, 0.6),
     "reset_opacity": m.reset_opacity,
     "fix_opacity": m.fix_opacity}[name]()
    # (fix_opacity / reset_opacity go through sigmoid and logit: the device's transcendental functions are within an
    #  ulp of the host's, not equal to them)
    check(m, d, opacity_tol=2e-6 if name in ("fix_opacity", "reset_opacity") else 0.0)
    if name in ("densify", "curvature"):
        assert m._curve_points.shape[0] > n_before
    for p in m.parameters():            # the optimizer still steps on the edited groups
        if p.requires_grad:
            p.grad = torch.ones_like(p)
    m.optimizer.step()


def test_mask_trim_split_on_device_matches_reference(cuda_dev):
    d = load("mask_trim")
    m = model_from(d, cuda_dev)
    m.mask_trim_split(0.7)
    check(m, d, mask_tol=1e-5)      # bilinear resampling of mask logits up to ~7: fp32 with the device's FMA contraction
