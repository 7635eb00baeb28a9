"""CPU, world_size 2 over gloo: view sharding + flat-gradient all-reduce reproduce the serial sum over all views."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from curve_gaussian_b200.parallel import FlatGrad, shard_views


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _view_loss(params, view_seed):
    """A stand-in per-view loss that touches every parameter differently per view."""
    g = torch.Generator().manual_seed(view_seed)
    return sum((p * torch.randn(p.shape, generator=g)).sum() + 0.5 * (p ** 2).sum() * (view_seed + 1) for p in params)


def _make_params():
    g = torch.Generator().manual_seed(0)
    shapes = [(7, 4, 3), (7, 1), (7, 1), (7, 5, 1)]   # curve_points, width, opacity, mask
    return [torch.randn(s, generator=g).requires_grad_(True) for s in shapes]


def _worker(rank, world, port, views, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    params = _make_params()
    fg = FlatGrad(params)
    fg.zero()
    for v in shard_views(views, rank, world):
        _view_loss(params, v).backward()
    flat = fg.all_reduce()
    if rank == 0:
        torch.save(flat.clone(), out)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_views_all_reduce_equals_serial_sum(tmp_path):
    views = list(range(7))   # ragged: 4 + 3
    out = str(tmp_path / "flat.pt")
    mp.spawn(_worker, args=(2, _free_port(), views, out), nprocs=2, join=True)
    params = _make_params()
    fg = FlatGrad(params)
    for v in views:
        _view_loss(params, v).backward()
    torch.testing.assert_close(torch.load(out), fg.flat, rtol=1e-6, atol=1e-6)


def test_shard_views_partition():
    views = list(range(10))
    parts = [shard_views(views, r, 4) for r in range(4)]
    assert sorted(sum(parts, [])) == views
    assert parts[0] == [0, 4, 8] and parts[3] == [3, 7]
    assert shard_views([], 0, 2) == []


def test_flat_grad_views_accumulate_in_place():
    params = _make_params()
    fg = FlatGrad(params)
    ptr = fg.flat.data_ptr()
    _view_loss(params, 1).backward()
    _view_loss(params, 2).backward()
    assert params[0].grad.data_ptr() == ptr          # still a view of the flat buffer
    assert fg.flat.abs().sum() > 0
    fg.zero()
    assert params[3].grad.abs().sum() == 0


def test_balanced_view_groups():
    from curve_gaussian_b200.parallel import balanced_view_groups
    costs = [5.0, 1.0, 9.0, 2.0, 8.0, 1.5, 4.0]
    groups = balanced_view_groups(costs, 2)
    assert groups == [[2, 4], [0, 6], [3, 5]]                     # sorted by cost, cut in pairs, remainder dropped
    flat = [i for g in groups for i in g]
    assert len(set(flat)) == len(flat)
    # the per-step maximum, summed over steps, never exceeds round-robin's on the same views
    rr = [costs[i:i + 2] for i in range(0, 6, 2)]
    assert sum(max(costs[i] for i in g) for g in groups) <= sum(max(g) for g in rr)
    assert balanced_view_groups(costs, 1) == [[2], [4], [0], [6], [3], [5], [1]]
