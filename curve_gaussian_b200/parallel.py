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


class FlatGrad:
    """One contiguous gradient buffer behind several parameters."""

    def __init__(self, params: Iterable[torch.Tensor]):
        self.params = [p for p in params]
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
