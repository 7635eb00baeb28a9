"""CPU: host logic of the sync-free binning policy (rasterizer.CapacityBinning) - no kernels involved."""
import pytest

from curve_gaussian_b200 import rasterizer as rz


def test_capacity_tracks_the_largest_count_with_headroom():
    pol = rz.CapacityBinning(headroom=1.5, granule=1000)
    key = (100, 64, 48)
    assert pol.capacity(key) is None           # unknown shape: the first forward takes the exact path
    pol.learn(key, 10_000)
    assert pol.capacity(key) == 15_000
    pol.learn(key, 4_000)                       # smaller counts never shrink it
    assert pol.capacity(key) == 15_000 and pol.max_seen[key] == 10_000
    pol.learn(key, 10_001)
    assert pol.capacity(key) == 16_000          # rounded up to the granule
    pol.learn((7, 8, 9), 1)
    assert pol.capacity((7, 8, 9)) == 1000      # at least one granule
    assert key in pol._static                   # the pinned slot graph replays write to exists before any capture


def test_accounting_raises_once_per_overflow_and_grows():
    pol = rz.CapacityBinning(headroom=1.25, granule=64)
    key = (10, 32, 32)
    pol.learn(key, 640)
    cap = pol.capacity(key)
    assert cap == 832
    assert pol._account(key, 700, 0, cap) is False and pol.capacity(key) == 896
    assert pol._account(key, 5000, 1, cap) is True
    assert pol.overflows == 1 and pol.capacity(key) >= 5000
    # slots written by graph replays are reported once
    slot = pol._static[key]
    slot[0], slot[1], slot[2] = 9000, 1, 6272
    with pytest.raises(rz.CapacityOverflow):
        pol.check()
    pol.check()
    assert pol.capacity(key) >= 9000


def test_switch_nests_and_restores():
    assert rz._policy is None
    with rz.capacity_binning(headroom=2.0) as a:
        assert rz._policy is a
        with rz.capacity_binning() as b:
            assert rz._policy is b and b is not a
        assert rz._policy is a
    assert rz._policy is None


def test_static_camera_takes_only_views_of_its_own_shape():
    import torch
    from curve_gaussian_b200 import synth
    from curve_gaussian_b200.graph import StaticCamera
    a, b = synth.random_cameras(2, 64, 48, seed=0)
    other = synth.random_cameras(1, 80, 48, seed=0)[0]
    sc = StaticCamera(a, device="cpu")
    ptrs = (sc.world_view_transform.data_ptr(), sc.full_proj_transform.data_ptr(), sc.camera_center.data_ptr())
    assert torch.equal(sc.world_view_transform, a.world_view_transform)
    sc.load(b)
    assert torch.equal(sc.full_proj_transform, b.full_proj_transform) and torch.equal(sc.camera_center, b.camera_center)
    assert ptrs == (sc.world_view_transform.data_ptr(), sc.full_proj_transform.data_ptr(), sc.camera_center.data_ptr())
    with pytest.raises(ValueError):
        sc.load(other)


def test_graphed_step_control_flow_with_a_fake_cuda(monkeypatch):
    """Host logic of GraphedStep.capture / verify (warm-up, capture, retry after a failed capture, re-capture after an
    overflow) with torch.cuda's stream / graph objects replaced by fakes - the real thing runs in the GPU tests."""
    import contextlib
    import torch
    from curve_gaussian_b200 import graph as G

    class FakeStream:
        def wait_stream(self, other): pass
        def synchronize(self): pass

    class FakeGraph:
        replays = 0
        def replay(self): FakeGraph.replays += 1

    fail_first = {"n": 1}

    @contextlib.contextmanager
    def fake_graph_ctx(g, stream=None):
        yield
        if fail_first["n"] > 0:
            fail_first["n"] -= 1
            raise RuntimeError("CUDA error: operation failed due to a previous error during capture")

    cur = FakeStream()
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a: cur)
    monkeypatch.setattr(torch.cuda, "Stream", FakeStream)
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "CUDAGraph", FakeGraph)
    monkeypatch.setattr(torch.cuda, "graph", fake_graph_ctx)
    restored = []
    monkeypatch.setattr(torch.cuda, "set_stream", restored.append)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)

    calls, prepared = [], []
    pol = rz.CapacityBinning(granule=64)
    gs = G.GraphedStep(lambda: calls.append(rz._policy) or "out", policy=pol,
                       calibrate=[lambda: prepared.append(1), lambda: prepared.append(2)])
    assert gs.replay() == "out"                                  # captures on first use
    assert prepared == [1, 2] and gs.captures == 1 and FakeGraph.replays == 1
    assert restored == [cur]                                     # the failed first capture put the stream back
    assert len(calls) == 2 + 2 + 1 + 1 + 1                       # calibrate, warm-up, failed capture, re-warm, capture
    assert all(c is pol for c in calls) and rz._policy is None   # policy active inside, restored outside
    assert gs.verify() and gs.captures == 1
    key = (10, 8, 8)
    pol.learn(key, 100)
    slot = pol._static[key]
    slot[0], slot[1], slot[2] = 500, 1, 128                      # what a replay that overflowed leaves behind
    assert not gs.verify() and gs.captures == 2 and pol.capacity(key) >= 500
    assert gs.verify()


def test_masked_mean_regularisers_equal_the_reference_formulations():
    """trainer.opacity_regulariser / width_regulariser against train.py:114-117 and :126-131 written with boolean
    indexing, values and gradients, including the empty cases."""
    import torch
    from curve_gaussian_b200.trainer import opacity_regulariser, width_regulariser
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(200, 1, generator=g, dtype=torch.float64, requires_grad=True)
    visible = torch.rand(200, generator=g) > 0.4
    ours = opacity_regulariser(torch.sigmoid(logits), visible)
    (ga,) = torch.autograd.grad(ours, logits)
    ref = torch.log(1 + torch.sigmoid(logits)[visible] ** 2 / 0.5).mean()
    (gb,) = torch.autograd.grad(ref, logits)
    assert torch.allclose(ours, ref, rtol=1e-12) and torch.allclose(ga, gb, rtol=1e-10, atol=1e-15)
    none = opacity_regulariser(torch.sigmoid(logits), torch.zeros(200, dtype=torch.bool))
    assert float(none) == 0.0 and float(torch.autograd.grad(none, logits)[0].abs().max()) == 0.0

    logw = (torch.randn(50, 1, generator=g, dtype=torch.float64) * 0.5 + torch.log(torch.tensor(5e-3))).requires_grad_(True)
    w = torch.exp(logw)
    ours = width_regulariser(w)
    (ga,) = torch.autograd.grad(ours, logw, retain_graph=True)
    mask = w >= 0.005
    assert 0 < int(mask.sum()) < 50
    ref = (w[mask] - 0.005).mean()
    (gb,) = torch.autograd.grad(ref, logw)
    assert torch.allclose(ours, ref, rtol=1e-12) and torch.allclose(ga, gb, rtol=1e-10, atol=1e-18)
    assert float(width_regulariser(torch.full((5, 1), 1e-3))) == 0.0


def test_balanced_view_partition_equal_sizes_and_near_equal_cost():
    import random
    from curve_gaussian_b200.parallel import balanced_view_partition
    rng = random.Random(3)
    for world, per in ((2, 256), (4, 128), (8, 64), (8, 1), (1, 7)):
        costs = [rng.uniform(6.0e6, 1.0e7) for _ in range(world * per)]
        parts = balanced_view_partition(costs, world)
        assert sorted(i for p in parts for i in p) == list(range(len(costs)))      # every view exactly once
        assert all(len(p) == per for p in parts)
        tot = [sum(costs[i] for i in p) for p in parts]
        assert max(tot) - min(tot) <= max(costs) - min(costs) + 1e-6                # within one view's cost spread
    import pytest
    with pytest.raises(ValueError):
        balanced_view_partition([1.0, 2.0, 3.0], 2)
