"""GPU parity of the non-rasterizer ops, called through the C ABI:
  sampling  vs golden vectors from the reference Python + vs the torch restatement at larger sizes
  fused SSIM vs the reference CUDA extension (oracle/_ref) and the CPU oracle
  3-NN dist vs the reference CUDA extension (oracle/_ref) and the CPU oracle
"""
import glob
import os

import numpy as np
import pytest
import torch

from curve_gaussian_b200 import sampling, synth
from curve_gaussian_b200.knn import distCUDA2
from curve_gaussian_b200.ssim import fused_ssim, _fusedssim, _fusedssim_backward
from oracle import cpu as oracle_cpu
from oracle import torch_ref
from tests import refload

pytestmark = pytest.mark.gpu
GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "sampling_*.npz")))


def rel(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_sampling_matches_reference_golden(cuda_dev, path):
    z = np.load(path)
    d = {k: torch.from_numpy(z[k]) for k in z.files if z[k].ndim}
    n = int(z["n"])
    cp = d["curve_points"].to(cuda_dev).requires_grad_(True)
    width = d["width"].to(cuda_dev).requires_grad_(True)
    xyz, rot, scal = sampling.sample_curves(cp, width, d["is_bezier"].to(cuda_dev), sampling.sample_t(n, cuda_dev))
    for name, o in (("xyz", xyz), ("rotation", rot), ("scaling", scal)):
        assert rel(o, d[name]) <= 1e-5, name
    # gradient: only the three sampling outputs take part here
    loss = (xyz * d["w_xyz"].to(cuda_dev)).sum() + (rot * d["w_rot"].to(cuda_dev)).sum() \
        + (scal * d["w_scal"].to(cuda_dev)).sum()
    loss.backward()
    cp2 = d["curve_points"].clone().requires_grad_(True)
    w2 = d["width"].clone().requires_grad_(True)
    x2, r2, s2 = torch_ref.sample_curves(cp2, w2, d["is_bezier"], n)
    ((x2 * d["w_xyz"]).sum() + (r2 * d["w_rot"]).sum() + (s2 * d["w_scal"]).sum()).backward()
    assert rel(cp.grad, cp2.grad) <= 2e-5
    assert rel(width.grad, w2.grad) <= 2e-5


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_sampling_and_activation_gradients_match_reference_golden(cuda_dev, path):
    """dL/d{control points, width, opacity, mask} of the CUDA sampling + activation ops against the gradients the
    REFERENCE's own autograd produced for the same weighted sum of all seven per-Gaussian outputs
    (tests/golden/make_sampling_golden.py: g_curve_points, g_width, g_opacity, g_mask), <= 1e-5."""
    from curve_gaussian_b200.activation import curve_activate
    from tests import parity
    z = np.load(path)
    d = {k: torch.from_numpy(z[k]) for k in z.files if z[k].ndim}
    n, use_mask, dev = int(z["n"]), bool(z["use_mask"]), cuda_dev
    leaf = lambda k: d[k].to(dev).requires_grad_(True)
    cp, width, opl, mask = leaf("curve_points"), leaf("width"), leaf("opacity_logit"), leaf("mask_logit")
    xyz, rot, scal = sampling.sample_curves(cp, width, d["is_bezier"].to(dev), sampling.sample_t(n, dev))
    rot_n, opacity, scales, all_map = curve_activate(xyz, rot, scal, opl, mask, n, d["cam_center"].to(dev),
                                                     d["world_view"].to(dev), use_mask, 0.01)
    outs = (xyz, rot, scal, opacity, rot_n, scales, all_map[:, :3])
    keys = ("xyz", "rot", "scal", "opacity", "rot_n", "scales", "local")
    for o, k, ref in zip(outs, keys, ("xyz", "rotation", "scaling", "opacity", "rot_normalized", "scales_masked", "local_axis")):
        assert rel(o, d[ref]) <= 1e-5, ref
    sum((o * d["w_" + k].to(dev)).sum() for o, k in zip(outs, keys)).backward()
    case = os.path.basename(path)
    for name, p_ in (("g_curve_points", cp), ("g_width", width), ("g_opacity", opl), ("g_mask", mask)):
        g = p_.grad if p_.grad is not None else torch.zeros_like(p_)
        e = rel(g, d[name]) if d[name].abs().max() > 0 else float(g.abs().max())
        parity.record("sampling_vs_reference_golden", case=case, quantity=name, max_rel=e, rel_err=parity.rel_err(g.cpu(), d[name]), tol=1e-5)
        assert e <= 1e-5, (name, e)


@pytest.mark.parametrize("B,n,lines", [(500, 12, 0.0), (300, 100, 0.3), (4000, 33, 0.5)])
def test_sampling_matches_torch_restatement(cuda_dev, B, n, lines):
    cp0, w0, _, isb = synth.random_curves(B, seed=B, line_fraction=lines)
    gen = torch.Generator().manual_seed(B)
    ws = [torch.randn(B * n, k, generator=gen) for k in (3, 4, 3)]
    cp = cp0.to(cuda_dev).requires_grad_(True)
    w = w0.to(cuda_dev).requires_grad_(True)
    outs = sampling.sample_curves(cp, w, isb.to(cuda_dev), sampling.sample_t(n, cuda_dev))
    sum((o * g.to(cuda_dev)).sum() for o, g in zip(outs, ws)).backward()
    # forward: against the fp32 restatement (same op order; |B(t)-B(t-h)| cancels ~2 digits in ANY fp32 evaluation)
    outs32 = torch_ref.sample_curves(cp0, w0, isb, n)
    for name, a, b in zip(("xyz", "rotation", "scaling"), outs, outs32):
        assert rel(a, b) <= 1e-5, name
    # gradients: against an fp64 evaluation of the same formulas
    cp2 = cp0.double().requires_grad_(True)
    w2 = w0.double().requires_grad_(True)
    outs2 = sample_curves_f64(cp2, w2, isb, n)
    sum((o * g.double()).sum() for o, g in zip(outs2, ws)).backward()
    assert rel(cp.grad, cp2.grad) <= 1e-4
    assert rel(w.grad, w2.grad) <= 1e-4


def sample_curves_f64(cp, w, isb, n):
    import oracle.torch_ref as tr
    old = tr.sample_t
    tr.sample_t = lambda n_, device="cpu": old(n_, device).double()
    try:
        return tr.sample_curves(cp, w, isb, n)
    finally:
        tr.sample_t = old


@pytest.mark.parametrize("shape", [(1, 1, 128, 128), (2, 3, 97, 211), (1, 1, 1080, 1920)])
def test_ssim_matches_reference_cuda(cuda_dev, shape):
    ref = refload.ref_ssim()
    if ref is None:
        pytest.skip("oracle/_ref/fused_ssim_cuda.so not built")
    g = torch.Generator().manual_seed(0)
    a = torch.rand(shape, generator=g).to(cuda_dev)
    b = torch.rand(shape, generator=g).to(cuda_dev)
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    m, d1, d2, d3 = _fusedssim(C1, C2, a, b, True)
    rm, r1, r2, r3 = ref.fusedssim(C1, C2, a, b, True)
    for x, y in ((m, rm), (d1, r1), (d2, r2), (d3, r3)):
        assert rel(x, y) <= 1e-5
    dL = torch.rand(shape, generator=g).to(cuda_dev)
    gi = _fusedssim_backward(C1, C2, a, b, dL, d1, d2, d3)
    gr = ref.fusedssim_backward(C1, C2, a, b, dL, r1, r2, r3)
    assert rel(gi, gr) <= 1e-5


def test_ssim_matches_cpu_oracle_and_autograd(cuda_dev):
    g = torch.Generator().manual_seed(1)
    a = torch.rand(1, 2, 70, 90, generator=g)
    b = torch.rand(1, 2, 70, 90, generator=g)
    m_o, *_ = oracle_cpu.ssim_fwd(a.numpy(), b.numpy())
    ad = a.to(cuda_dev).requires_grad_(True)
    val = fused_ssim(ad, b.to(cuda_dev))
    assert abs(val.item() - float(m_o.mean())) <= 1e-5
    val.backward()
    d = oracle_cpu.ssim_fwd(a.numpy(), b.numpy())
    go = oracle_cpu.ssim_bwd(a.numpy(), b.numpy(), np.full(a.shape, 1.0 / a.numel(), np.float32), d[1], d[2], d[3])
    assert rel(ad.grad, torch.from_numpy(go)) <= 1e-5
    # "valid" padding path
    v = fused_ssim(a.to(cuda_dev), b.to(cuda_dev), padding="valid", train=False)
    assert abs(v.item() - float(m_o[:, :, 5:-5, 5:-5].mean())) <= 1e-5


@pytest.mark.parametrize("P", [4, 100, 3375, 50000])
def test_knn_matches_reference_and_oracle(cuda_dev, P):
    g = torch.Generator().manual_seed(P)
    if P == 3375:   # the reference's 15^3 init grid (scene/dataset_readers.py:404-412)
        ax = torch.linspace(0, 1, 15)
        pts = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(-1, 3).contiguous()
    else:
        pts = torch.rand(P, 3, generator=g)
    out = distCUDA2(pts.to(cuda_dev))
    if P <= 5000:
        o = oracle_cpu.knn_mean_dist2(pts.numpy())
        assert np.array_equal(out.cpu().numpy(), o)
    ref = refload.ref_knn()
    if ref is not None:
        r = ref.distCUDA2(pts.to(cuda_dev))
        assert torch.equal(out, r)
