"""View-parallel host logic (one process per GPU, torch.distributed).

Views are independent given replicated curve parameters: rank r renders views
r::world and the curve-parameter gradients of all ranks are summed with ONE all-reduce per step.
The parameters' .grad tensors are views into a single flat fp32 buffer, so backward accumulates
straight into the buffer the collective reduces (no pack / unpack copies).
"""
from __future__ import annotations

from typing import Iterable, List, Sequence

import torch
import torch.distributed as dist


def shard_views(views: Sequence, rank: int, world: int) -> List:
    """Round-robin view assignment: rank r gets views r, r+world, ..."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return list(views[rank::world])


def balanced_view_groups(costs: Sequence[float], world: int) -> List[List[int]]:
    """Group view indices so that the `world` views rendered in the same step cost about the same.

    Every step ends with an all-reduce, so a step takes as long as its most expensive view; with views of
    uneven cost (the number of tile-instances R varies by tens of percent between cameras) round-robin
    sharding pays max-over-ranks every step. Sorting by cost and cutting the sorted list into consecutive
    groups of `world` makes the members of a group near-equal. Returns groups[s][r] = index of the view rank r
    renders in step s; views that do not fill a last complete group are dropped (as round-robin would leave
    some ranks idle there)."""
    if world < 1:
        raise ValueError("world must be >= 1")
    order = sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i))
    return [order[s * world:(s + 1) * world] for s in range(len(order) // world)]


def balanced_view_partition(costs: Sequence[float], world: int) -> List[List[int]]:
    """Partition ALL view indices into `world` shards of equal size and near-equal total cost.

    For steps in which every rank accumulates several views before the one all-reduce (BASELINE.json configs[4]:
    512 views over 8 GPUs, 64 per rank): the step lasts as long as the rank whose views cost most in total. Views are
    taken by descending cost and dealt to the ranks boustrophedon (0..w-1, w-1..0, ...), which keeps the shard
    sizes equal and the totals within one view's cost of each other. Returns parts[r] = view indices of rank r;
    len(costs) must be a multiple of world."""
    if world < 1:
        raise ValueError("world must be >= 1")
    if len(costs) % world:
        raise ValueError(f"{len(costs)} views do not split evenly over {world} ranks")
    order = sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i))
    parts: List[List[int]] = [[] for _ in range(world)]
    for k, i in enumerate(order):
        rnd, pos = divmod(k, world)
        parts[pos if rnd % 2 == 0 else world - 1 - pos].append(i)
    return parts


class FlatGrad:
    """One contiguous gradient buffer behind several parameters.

    direct=True additionally lets the library's backward kernels ADD their curve-parameter gradients straight into
    this buffer (sampling.py / activation.py check the flag): no AccumulateGrad node runs for these parameters, so a
    step costs neither the ATen add / fill kernels nor - inside a CUDA-graph capture - a node bound to another stream.
    The buffer must then be zeroed by the caller before every accumulation round (zero()), as with any .grad."""

    def __init__(self, params: Iterable[torch.Tensor], direct: bool = False):
        self.params = [p for p in params]
        self.direct = bool(direct)
        if not self.params:
            raise ValueError("no parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        for p in self.params:
            if p.device != dev or p.dtype != dt:
                raise ValueError("parameters must share device and dtype")
        self.flat = torch.zeros(sum(p.numel() for p in self.params), device=dev, dtype=dt)
        self.bind()

    def bind(self) -> None:
        """(Re)attach .grad views; call again after an optimizer step that replaced .grad."""
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            p._cg_direct_grad = self.direct
            off += p.numel()

    def zero(self) -> None:
        self.flat.zero_()

    def all_reduce(self, average: bool = False):
        """Sum (or average) the flat buffer over all ranks; no-op without a process group."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            if average:
                self.flat.div_(dist.get_world_size())
        return self.flat
