"""GPU: sync-free binning (cg_raster_fwd_capacity, rasterizer.CapacityBinning) against the exact
two-stage forward, overflow reporting, and CUDA-graph replay of a whole step (graph.GraphedStep) against the
eager step. Integer state and pixels must be bit-identical (same kernels, same order); gradients differ only
by the order of fp32 atomics (tolerance 1e-5 max-rel, the reference's own run-to-run noise)."""
import math

import pytest
import torch

from curve_gaussian_b200 import _lib, synth
from curve_gaussian_b200 import rasterizer as rz
from curve_gaussian_b200.curve_model import GaussianCurveModel
from curve_gaussian_b200.graph import GraphedStep, StaticCamera
from curve_gaussian_b200.loss import edge_ssim_loss
from curve_gaussian_b200.parallel import FlatGrad
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


def scene(dev, B=300, n=20, W=320, H=208, seed=3):
    cp, width, opl, isb = synth.random_curves(B, seed=seed, line_fraction=0.2)
    width = width + 0.5
    cam = synth.random_cameras(1, W, H, seed=seed + 1)[0].to(dev)
    xyz, q, sc = torch_ref.sample_curves(cp.to(dev), width.to(dev), isb.to(dev), n)
    mask = torch.ones(B, n, 1, device=dev)
    m3, op, scl, rot, col, amap = torch_ref.raster_inputs(xyz, q, sc, opl.to(dev), n, mask, cam.camera_center,
                                                          cam.world_view_transform)
    rs = rz.GaussianRasterizationSettings(H, W, math.tan(cam.FoVx / 2), math.tan(cam.FoVy / 2),
                                          torch.full((3,), 0.1, device=dev), 1.0, cam.world_view_transform,
                                          cam.full_proj_transform, 0, cam.camera_center, False, False, False, True)
    return rs, [t.detach().contiguous() for t in (m3, col, op, scl, rot, amap)]


def fetch(which, P, R, rs, geom, img, bk, scratch, n, dtype):
    lib = _lib.load()
    dst = torch.empty(n, dtype=dtype, device=geom.device)
    _lib.check(lib.cg_raster_debug_fetch(which, P, R, rs.image_width, rs.image_height, geom.data_ptr(), img.data_ptr(),
                                         bk.data_ptr(), _lib.ptr(scratch), dst.data_ptr(), _lib.stream(geom.device)),
               "debug_fetch")
    return dst


@pytest.mark.parametrize("slack", [0, 1, 4097, 100000])
def test_capacity_forward_and_backward_match_exact_path(cuda_dev, slack):
    dev = cuda_dev
    rs, (m3, col, op, scl, rot, amap) = scene(dev)
    P, W, H = m3.shape[0], rs.image_width, rs.image_height
    R, color, radii, geom, bk, img, invd, omap = rz.rasterize_forward_raw(rs, m3, col, op, scl, rot, None, amap)
    assert R > 1000
    cap = R + slack
    R2, color2, radii2, geom2, bk2, img2, invd2, omap2 = rz.rasterize_forward_raw(rs, m3, col, op, scl, rot, None, amap,
                                                                                   capacity=cap)
    scratch2 = rz.rasterize_forward_raw.last_scratch
    counter = rz.rasterize_forward_raw.last_counter.cpu()
    assert R2 == cap and counter.tolist() == [R, 0]
    for a, b in ((color, color2), (radii, radii2), (invd, invd2), (omap, omap2)):
        assert torch.equal(a, b)
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    for which, n, dt in ((7, W * H, torch.int32), (8, W * H, torch.float32), (2, tiles * 2, torch.int32)):
        assert torch.equal(fetch(which, P, R, rs, geom, img, bk, None, n, dt),
                           fetch(which, P, cap, rs, geom2, img2, bk2, None, n, dt)), which
    pl1 = fetch(1, P, R, rs, geom, img, bk, None, R, torch.int32)
    pl2 = fetch(1, P, cap, rs, geom2, img2, bk2, None, cap, torch.int32)[:R]
    assert torch.equal(pl1, pl2)

    g = torch.Generator().manual_seed(0)
    gc, gi, gm = (torch.randn(s, generator=g).to(dev) for s in ((1, H, W), (1, H, W), (4, H, W)))
    a = rz.rasterize_backward_raw(rs, m3, radii, col, amap, op, scl, rot, None, gc, gi, gm, geom, R, bk, img)
    b = rz.rasterize_backward_raw(rs, m3, radii2, col, amap, op, scl, rot, None, gc, gi, gm, geom2, cap, bk2, img2)
    for i, (x, y) in enumerate(zip(a, b)):
        if x is None or x.numel() == 0:
            continue
        assert rel(y, x) <= 1e-5, i


def test_capacity_overflow_is_reported_and_the_policy_grows(cuda_dev):
    dev = cuda_dev
    rs, (m3, col, op, scl, rot, amap) = scene(dev, seed=5)
    P, W, H = m3.shape[0], rs.image_width, rs.image_height
    R = rz.rasterize_forward_raw(rs, m3, col, op, scl, rot, None, amap)[0]
    small = R // 2
    out = rz.rasterize_forward_raw(rs, m3, col, op, scl, rot, None, amap, capacity=small)
    assert rz.rasterize_forward_raw.last_counter.cpu().tolist() == [R, 1]
    assert torch.isfinite(out[1]).all()       # truncated list: a wrong image, never a crash
    # the point list holds the first `small` instances of the exact emission (depth) order, sorted by tile
    with rz.capacity_binning(headroom=1.25, granule=256) as pol:
        key = (P, W, H)
        pol.caps[key] = small                  # a stale, too small capacity
        pol.max_seen[key] = small
        pol._static[key] = torch.zeros(3, dtype=torch.int32).pin_memory()
        rz.rasterize_forward_raw(rs, m3, col, op, scl, rot, None, amap)
        torch.cuda.synchronize()
        with pytest.raises(rz.CapacityOverflow):
            pol.poll()
        assert pol.caps[key] >= R and pol.overflows == 1
        out2 = rz.rasterize_forward_raw(rs, m3, col, op, scl, rot, None, amap)
        torch.cuda.synchronize()
        pol.poll()                             # fits now
    assert out2[0] == pol.caps[key]
    ref = rz.rasterize_forward_raw(rs, m3, col, op, scl, rot, None, amap)
    assert ref[0] == R and torch.equal(ref[1], out2[1])


def test_zero_instances_in_capacity_mode(cuda_dev):
    dev = cuda_dev
    rs, (m3, col, op, scl, rot, amap) = scene(dev, B=20, n=8)
    behind = m3 - 50.0 * rs.viewmatrix[:3, 2]   # everything behind the camera: R == 0
    out = rz.rasterize_forward_raw(rs, behind, col, op, scl, rot, None, amap, capacity=4096)
    assert rz.rasterize_forward_raw.last_counter.cpu().tolist() == [0, 0]
    assert torch.all(out[1] == 0.1) and int(out[2].abs().max()) == 0
    g = rz.rasterize_backward_raw(rs, behind, out[2], col, amap, op, scl, rot, None, torch.ones_like(out[1]), None, None,
                                  out[3], 4096, out[4], out[5])
    assert float(g[3].abs().max()) == 0


def test_graphed_step_replays_match_eager_steps(cuda_dev):
    dev = cuda_dev
    B, n, W, H = 250, 16, 256, 192
    cp, width, opl, isb = synth.random_curves(B, seed=11, line_fraction=0.25)
    width = width + 0.6
    cams = [c.to(dev) for c in synth.random_cameras(3, W, H, seed=12)]
    g = torch.Generator().manual_seed(4)
    gts = [torch.rand(1, H, W, generator=g).to(dev) for _ in cams]
    bg = torch.zeros(3, device=dev)

    def make(direct=False):
        m = GaussianCurveModel(0, n_gaussians=n, device=dev).create_from_curves(cp, width, opl, isb)
        return m, FlatGrad([m._curve_points, m._width, m._opacity, m._mask], direct=direct)

    # eager, exact path
    model, fg = make()
    eager = []
    for cam, gt in zip(cams, gts):
        fg.zero()
        model.prepare_scaling_rot()
        loss = edge_ssim_loss(render(cam, model, Pipe(), bg)["render_raw"], gt, clamp=True)
        loss.backward()
        eager.append((loss.detach().clone(), fg.flat.clone()))

    # one captured step, replayed per view; the kernels add the curve gradients straight into the flat buffer, so the
    # captured backward contains no AccumulateGrad node (the node that, bound to the stream of an earlier eager step,
    # made this capture fail under compute-sanitizer in round 1: profiles/r01_memcheck_capacity.txt)
    model2, fg2 = make(direct=True)
    scam = StaticCamera(cams[0])
    gt_static = torch.empty_like(gts[0])

    def step():
        fg2.zero()
        model2.prepare_scaling_rot()
        loss = edge_ssim_loss(render(scam, model2, Pipe(), bg)["render_raw"], gt_static, clamp=True)
        loss.backward()
        return loss

    gs = GraphedStep(step, calibrate=[(lambda c=c: scam.load(c)) for c in cams])
    gt_static.copy_(gts[0])
    gs.capture()
    for rep in range(2):
        for i, (cam, gt) in enumerate(zip(cams, gts)):
            scam.load(cam)
            gt_static.copy_(gt)
            loss = gs.replay()
            torch.cuda.synchronize()
            assert gs.verify()
            assert torch.equal(loss.detach(), eager[i][0]), (rep, i)
            assert rel(fg2.flat, eager[i][1]) <= 1e-5, (rep, i)
    assert gs.captures == 1


def test_graphed_step_recaptures_after_overflow(cuda_dev):
    dev = cuda_dev
    B, n, W, H = 150, 16, 192, 128
    cp, width, opl, isb = synth.random_curves(B, seed=21)
    width = width + 0.6
    cams = [c.to(dev) for c in synth.random_cameras(8, W, H, seed=22)]
    bg = torch.zeros(3, device=dev)
    model = GaussianCurveModel(0, n_gaussians=n, device=dev).create_from_curves(cp, width, opl, isb)
    model.prepare_scaling_rot()
    refs, Rs = [], []
    with torch.no_grad():
        for c in cams:
            refs.append(render(c, model, Pipe(), bg)["render_raw"].clone())
            Rs.append(rz.rasterize_forward_raw.last_R)
    lo, hi = Rs.index(min(Rs)), Rs.index(max(Rs))
    assert Rs[hi] > Rs[lo] + 64
    scam = StaticCamera(cams[lo])

    def step():
        model.prepare_scaling_rot()
        with torch.no_grad():
            return render(scam, model, Pipe(), bg)["render_raw"] * 1.0

    pol = rz.CapacityBinning(headroom=1.0, granule=64)     # capacity = R of the cheapest view, no slack
    gs = GraphedStep(step, policy=pol).capture()
    assert torch.equal(gs.replay(), refs[lo])
    torch.cuda.synchronize()
    assert gs.verify()
    scam.load(cams[hi])                                     # more instances than the captured capacity
    gs.replay()
    torch.cuda.synchronize()
    assert not gs.verify() and gs.captures == 2             # overflow seen, step re-captured with a larger capacity
    img = gs.replay().clone()
    torch.cuda.synchronize()
    assert gs.verify()
    assert torch.equal(img, refs[hi])


def test_overflow_is_sticky_across_replays(cuda_dev):
    """An overflowing replay followed by a cheap one must still be reported: the {R, overflow} slot is a running
    maximum over the replays since the last check(), not the last replay's value."""
    dev = cuda_dev
    B, n, W, H = 150, 16, 192, 128
    cp, width, opl, isb = synth.random_curves(B, seed=21)
    width = width + 0.6
    cams = [c.to(dev) for c in synth.random_cameras(8, W, H, seed=22)]
    bg = torch.zeros(3, device=dev)
    model = GaussianCurveModel(0, n_gaussians=n, device=dev).create_from_curves(cp, width, opl, isb)
    model.prepare_scaling_rot()
    Rs = []
    with torch.no_grad():
        for c in cams:
            render(c, model, Pipe(), bg)
            Rs.append(rz.rasterize_forward_raw.last_R)
    lo, hi = Rs.index(min(Rs)), Rs.index(max(Rs))
    scam = StaticCamera(cams[lo])

    def step():
        model.prepare_scaling_rot()
        with torch.no_grad():
            return render(scam, model, Pipe(), bg)["render_raw"] * 1.0

    pol = rz.CapacityBinning(headroom=1.0, granule=64)
    gs = GraphedStep(step, policy=pol).capture()
    gs.replay()
    scam.load(cams[hi])      # overflows ...
    gs.replay()
    scam.load(cams[lo])      # ... and the next replay fits again
    gs.replay()
    torch.cuda.synchronize()
    assert not gs.verify(), "the overflow of the middle replay was lost"
    assert pol.max_seen[(B * n, W, H)] >= Rs[hi]


def test_multi_view_step_equals_serial_graph_replays(cuda_dev):
    """MultiViewStep: V captured copies of the step replayed concurrently. Per view the loss is bit-identical to a
    single GraphedStep replay of that view; the summed curve gradient equals the serial sum."""
    from curve_gaussian_b200.graph import MultiViewStep
    dev = cuda_dev
    B, n, W, H, V = 200, 12, 256, 192, 4
    cp, width, opl, isb = synth.random_curves(B, seed=41, line_fraction=0.2)
    width = width + 0.6
    cams = [c.to(dev) for c in synth.random_cameras(V, W, H, seed=42)]
    g = torch.Generator().manual_seed(6)
    gts = [torch.rand(1, H, W, generator=g).to(dev) for _ in cams]
    bg = torch.zeros(3, device=dev)

    # serial reference: eager steps, gradients summed in one flat buffer
    m0 = GaussianCurveModel(0, n_gaussians=n, device=dev).create_from_curves(cp, width, opl, isb)
    fg0 = FlatGrad([m0._curve_points, m0._width, m0._opacity, m0._mask])
    fg0.zero()
    losses0 = []
    for cam, gt in zip(cams, gts):
        m0.prepare_scaling_rot()
        loss = edge_ssim_loss(render(cam, m0, Pipe(), bg)["render_raw"], gt, clamp=True)
        loss.backward()
        losses0.append(loss.detach().clone())

    m1 = GaussianCurveModel(0, n_gaussians=n, device=dev).create_from_curves(cp, width, opl, isb)

    def body(scam, gt):
        m1.prepare_scaling_rot()
        loss = edge_ssim_loss(render(scam, m1, Pipe(), bg)["render_raw"], gt, clamp=True)
        loss.backward()
        return loss

    mv = MultiViewStep([m1._curve_points, m1._width, m1._opacity, m1._mask], body, cams[0], gts[0], V).capture(cams)
    for rep in range(2):
        outs = mv.replay(cams, gts)
        torch.cuda.synchronize()
        assert mv.verify()
        for k in range(V):
            assert torch.equal(outs[k].detach(), losses0[k]), (rep, k)
        assert rel(mv.total, fg0.flat) <= 1e-5, rep
        assert m1._curve_points.grad.data_ptr() == mv.total.data_ptr()
