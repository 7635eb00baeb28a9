"""Two ranks over NCCL (needs 2 GPUs: `gpurun --gpus 2`; skipped on a one-GPU box): the all-reduced flat curve
gradient of a step in which each rank accumulates its shard of the views equals the serial sum of the same views on
one GPU (BASELINE.json configs[4] semantics: views sharded, one all-reduce per step)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from curve_gaussian_b200 import synth
    from curve_gaussian_b200.curve_model import GaussianCurveModel
    from curve_gaussian_b200.loss import edge_ssim_loss
    from curve_gaussian_b200.parallel import FlatGrad, balanced_view_partition
    from curve_gaussian_b200.renderer import render
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    class Pipe:
        debug = False
        antialiasing = False
        render_geo = True

    B, n, W, H, NV = 300, 16, 320, 240, 8
    cp, width, opl, isb = synth.random_curves(B, seed=5)
    width = width + 0.5
    cams = synth.random_cameras(NV, W, H, seed=6)
    gts = [(torch.rand(1, H, W, generator=torch.Generator().manual_seed(100 + i)) > 0.9).float() for i in range(NV)]
    model = GaussianCurveModel(0, n_gaussians=n, device=dev).create_from_curves(cp, width, opl, isb)
    fg = FlatGrad([model._curve_points, model._width, model._opacity, model._mask])
    bg = torch.zeros(3, device=dev)

    def accumulate(views):
        fg.zero()
        for i in views:
            model.prepare_scaling_rot()
            pkg = render(cams[i].to(dev), model, Pipe(), bg)
            edge_ssim_loss(pkg["render_raw"], gts[i].to(dev), clamp=True).backward()

    part = balanced_view_partition([1.0 + (i % 3) for i in range(NV)], world)
    accumulate(part[rank])
    fg.all_reduce()
    sharded = fg.flat.clone()
    accumulate(range(NV))                      # the same views, serially, on this GPU alone
    serial = fg.flat.clone()
    torch.cuda.synchronize()
    err = ((sharded - serial).abs().max() / serial.abs().max()).item()
    q.put((rank, err, float(serial.abs().max())))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_views_all_reduce_equals_serial_sum():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29731
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, mx in res:
        assert mx > 0 and err <= 1e-5, (rank, err)
