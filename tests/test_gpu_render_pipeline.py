"""GPU: the drop-in surface end to end (GaussianCurveModel.prepare_scaling_rot -> render -> loss -> backward)
against the torch restatement of the reference's Python glue around the same CUDA rasterizer, and the fused
activation op against the unfused reference formulation."""
import math

import pytest
import torch

from curve_gaussian_b200 import synth
from curve_gaussian_b200.activation import curve_activate
from curve_gaussian_b200.curve_model import GaussianCurveModel
from curve_gaussian_b200.renderer import render
from oracle import torch_ref

pytestmark = pytest.mark.gpu


class Pipe:
    debug = False
    antialiasing = False
    render_geo = True


def rel(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


@pytest.mark.parametrize("use_mask", [False, True])
def test_fused_activation_matches_unfused_reference_formulation(cuda_dev, use_mask):
    dev = cuda_dev
    B, n = 300, 16
    cp, width, opl, isb = synth.random_curves(B, seed=5, line_fraction=0.2)
    g = torch.Generator().manual_seed(2)
    opl = opl + torch.randn(B, 1, generator=g)
    mask = torch.randn(B, n, 1, generator=g) * 3
    cam = synth.random_cameras(1, 320, 240, seed=4)[0].to(dev)
    xyz, rot, scal = torch_ref.sample_curves(cp, width, isb, n)
    leaves = [t.to(dev).requires_grad_(True) for t in (rot, scal, opl, mask)]
    ws = [torch.randn(B * n, k, generator=g).to(dev) for k in (4, 1, 3, 4)]
    outs = curve_activate(xyz.to(dev), leaves[0], leaves[1], leaves[2], leaves[3], n, cam.camera_center,
                          cam.world_view_transform, use_mask, 0.01)
    sum((o * w).sum() for o, w in zip(outs, ws)).backward()
    leaves2 = [t.to(dev).requires_grad_(True) for t in (rot, scal, opl, mask)]
    _, op2, sc2, rn2, _, am2 = torch_ref.raster_inputs(xyz.to(dev), leaves2[0], leaves2[1], leaves2[2], n, leaves2[3],
                                                      cam.camera_center, cam.world_view_transform, use_mask, 0.01)
    outs2 = (rn2, op2, sc2, am2)
    sum((o * w).sum() for o, w in zip(outs2, ws)).backward()
    for name, a, b in zip(("rot_n", "opacity", "scales", "all_map"), outs, outs2):
        assert rel(a, b) <= 1e-6, name
    for name, a, b in zip(("rotation", "scaling", "opacity_logit", "mask_logit"), leaves, leaves2):
        if b.grad is None:
            assert a.grad is None or a.grad.abs().max() == 0, name
            continue
        assert rel(a.grad, b.grad) <= 2e-5, name


@pytest.mark.parametrize("use_mask", [False, True])
def test_render_backward_to_curve_parameters(cuda_dev, use_mask):
    dev = cuda_dev
    B, n, W, H = 200, 24, 200, 150
    cp, width, opl, isb = synth.random_curves(B, seed=9, line_fraction=0.3)
    width = width + 0.7
    cam = synth.random_cameras(1, W, H, seed=8)[0].to(dev)
    g = torch.Generator().manual_seed(1)
    mask = torch.randn(B, n, 1, generator=g) * 2 + 2
    gimg = torch.randn(1, H, W, generator=g).to(dev)
    bg = torch.zeros(3, device=dev)

    model = GaussianCurveModel(0, n_gaussians=n, device=dev).create_from_curves(cp, width, opl, isb, mask)
    model.prepare_scaling_rot()
    pkg = render(cam, model, Pipe(), bg, use_mask=use_mask)
    (pkg["render"] * gimg).sum().backward()

    # same pipeline with the reference's unfused Python glue (torch autograd) around the same rasterizer
    ref = GaussianCurveModel(0, n_gaussians=n, device=dev).create_from_curves(cp, width, opl, isb, mask)
    ref.fuse_activations = False
    xyz, rot, scal = torch_ref.sample_curves(ref._curve_points, ref._width, ref.is_bezier, n)
    ref._xyz, ref._rotation, ref._scaling = xyz, rot, scal
    pkg2 = render(cam, ref, Pipe(), bg, use_mask=use_mask)
    (pkg2["render"] * gimg).sum().backward()

    assert torch.equal(pkg["radii"], pkg2["radii"])
    assert rel(pkg["render"], pkg2["render"]) <= 1e-5
    assert rel(pkg["rend_alpha"], pkg2["rend_alpha"]) <= 1e-5
    assert rel(pkg["rend_dir"], pkg2["rend_dir"]) <= 1e-4
    assert pkg["viewspace_points"].grad is not None
    assert rel(pkg["viewspace_points"].grad, pkg2["viewspace_points"].grad) <= 1e-5
    for name in ("_curve_points", "_width", "_opacity", "_mask"):
        a, b = getattr(model, name).grad, getattr(ref, name).grad
        if b is None or b.abs().max() == 0:
            continue
        assert rel(a, b) <= 1e-4, name


def test_empty_and_degenerate_inputs(cuda_dev):
    from curve_gaussian_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer
    dev = cuda_dev
    cam = synth.random_cameras(1, 64, 48, seed=1)[0].to(dev)
    rs = GaussianRasterizationSettings(48, 64, math.tan(cam.FoVx / 2), math.tan(cam.FoVy / 2), torch.zeros(3, device=dev),
                                       1.0, cam.world_view_transform, cam.full_proj_transform, 0, cam.camera_center,
                                       False, False, False, True)
    r = GaussianRasterizer(rs)
    z = lambda *s: torch.zeros(*s, device=dev)
    color, radii, invd, amap = r(means3D=z(0, 3), means2D=z(0, 3), opacities=z(0, 1), colors_precomp=z(0, 1),
                                 scales=z(0, 3), rotations=z(0, 4), all_map=z(0, 4))
    assert color.shape == (1, 48, 64) and float(color.abs().max()) == 0 and radii.numel() == 0
    # everything behind the camera: R == 0, image == background
    rs2 = rs._replace(bg=torch.full((3,), 0.25, device=dev))
    m = (cam.camera_center + torch.tensor([0.0, 0.0, 0.0], device=dev)).repeat(5, 1) \
        - 3.0 * cam.world_view_transform[:3, 2]
    color, radii, _, _ = GaussianRasterizer(rs2)(means3D=m, means2D=z(5, 3), opacities=z(5, 1) + 0.5,
                                                 colors_precomp=z(5, 1) + 1, scales=z(5, 3) + 0.01,
                                                 rotations=torch.tensor([[1.0, 0, 0, 0]], device=dev).repeat(5, 1),
                                                 all_map=z(5, 4))
    assert int(radii.abs().max()) == 0 and torch.all(color == 0.25)
    with pytest.raises(Exception):
        r(means3D=m, means2D=z(5, 3), opacities=z(5, 1))
    vis = r.markVisible(m)
    assert vis.dtype == torch.bool and not bool(vis.any())


def test_direct_gradient_accumulation_equals_autograd(cuda_dev):
    """FlatGrad(direct=True): the sampling / activation backward kernels add their curve-parameter gradients straight
    into the flat buffer. Same numbers as the AccumulateGrad path, over two accumulated views, mask included."""
    from curve_gaussian_b200.loss import edge_ssim_loss
    from curve_gaussian_b200.parallel import FlatGrad
    dev = cuda_dev
    B, n, W, H = 120, 16, 256, 192
    cp, width, opl, isb = synth.random_curves(B, seed=31, line_fraction=0.3)
    width = width + 0.6
    mask = torch.randn(B, n, 1, generator=torch.Generator().manual_seed(8)) * 3
    cams = [c.to(dev) for c in synth.random_cameras(2, W, H, seed=32)]
    gts = [(torch.rand(1, H, W, generator=torch.Generator().manual_seed(40 + i)) > 0.9).float().to(dev) for i in range(2)]
    bg = torch.zeros(3, device=dev)
    flats = []
    for direct in (False, True):
        m = GaussianCurveModel(0, n_gaussians=n, device=dev).create_from_curves(cp, width, opl, isb, mask)
        fg = FlatGrad([m._curve_points, m._width, m._opacity, m._mask], direct=direct)
        fg.zero()
        for cam, gt in zip(cams, gts):
            m.prepare_scaling_rot()
            pkg = render(cam, m, Pipe(), bg, use_mask=True, mask_thr=0.01)
            edge_ssim_loss(pkg["render_raw"], gt, clamp=True).backward()
        torch.cuda.synchronize()
        assert all(p.grad.data_ptr() >= fg.flat.data_ptr() for p in fg.params), "a .grad was replaced"
        flats.append(fg.flat.clone())
    assert flats[0].abs().max() > 0
    err = ((flats[0] - flats[1]).abs().max() / flats[0].abs().max()).item()
    assert err <= 1e-6, err    # only the rasterizer's atomics order differs between the two runs
