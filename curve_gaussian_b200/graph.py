"""CUDA-graph capture of one training step.

At the reference's own operating points (5 k - 100 k curve-Gaussians, images up to 1200x680) one step is
0.3 - 0.6 ms of kernels behind ~0.8 ms of host work: autograd bookkeeping, ~35 allocations, ~45 launches and
the read-back of the tile-instance count R in every forward (rasterizer_impl.cu:283-291). With sync-free
binning (`rasterizer.CapacityBinning`, `cg_raster_fwd_capacity`) nothing in

    sample -> activate -> rasterize -> loss -> backward

touches the host any more, so the whole step is captured once and replayed with one launch per step.
What changes between replays lives in fixed device tensors: the camera (`StaticCamera`), the target image
and the parameters themselves; gradients land in the parameters' existing `.grad` (e.g. the views of
`parallel.FlatGrad`).
"""
from __future__ import annotations

import gc
from typing import Callable, Optional, Sequence

import torch

from . import rasterizer as _rz


class StaticCamera:
    """A camera whose matrices live in fixed device tensors.

    Has the attributes `render()` reads from a reference `Camera` (scene/cameras.py:17-71: image size, FoV,
    world_view_transform, full_proj_transform, camera_center). `load(cam)` copies another view's matrices in;
    image size and field of view are baked into a captured step (they are kernel arguments), so every view
    loaded into one StaticCamera must share them - as the views of one dataset do.
    """

    def __init__(self, like, device=None):
        dev = torch.device(device) if device is not None else like.world_view_transform.device
        self.image_width = int(like.image_width)
        self.image_height = int(like.image_height)
        self.FoVx = float(like.FoVx)
        self.FoVy = float(like.FoVy)
        self.image_name = getattr(like, "image_name", "static")
        self.world_view_transform = torch.empty(4, 4, dtype=torch.float32, device=dev)
        self.full_proj_transform = torch.empty(4, 4, dtype=torch.float32, device=dev)
        self.camera_center = torch.empty(3, dtype=torch.float32, device=dev)
        self.load(like)

    def load(self, cam) -> "StaticCamera":
        if (int(cam.image_width), int(cam.image_height)) != (self.image_width, self.image_height) or \
                abs(float(cam.FoVx) - self.FoVx) > 1e-12 or abs(float(cam.FoVy) - self.FoVy) > 1e-12:
            raise ValueError("a StaticCamera only takes views with its own image size and field of view")
        self.world_view_transform.copy_(cam.world_view_transform, non_blocking=True)
        self.full_proj_transform.copy_(cam.full_proj_transform, non_blocking=True)
        self.camera_center.copy_(cam.camera_center, non_blocking=True)
        return self


class GraphedStep:
    """Capture `fn()` - one forward + backward - into a CUDA graph and replay it.

    `fn` must read everything that varies between steps from fixed device tensors and return a tensor (the
    loss) or a tuple of tensors; it must not synchronise, which for the rasterizer means a `CapacityBinning`
    policy has to be active (one is created if none is given). `calibrate` is an optional list of callables
    run before the capture, each followed by one eager `fn()`, to show the policy the largest R it will meet
    (e.g. one per distinct view: `lambda: scam.load(cam)`).

    Autograd graphs over the same parameters that were built on another stream must be gone when the capture
    starts (their AccumulateGrad nodes would tie the captured backward to that stream); `fn` itself should
    therefore rebuild everything it differentiates through, as `GaussianCurveModel.prepare_scaling_rot()` does
    by dropping the previous sampled tensors first.
    """

    def __init__(self, fn: Callable[[], object], policy: Optional[_rz.CapacityBinning] = None, warmup: int = 2,
                 calibrate: Optional[Sequence[Callable[[], None]]] = None):
        self.fn = fn
        self.policy = policy or _rz.CapacityBinning()
        self.warmup = max(2, int(warmup))   # the first eager run of a shape takes the exact path; the second the capacity one
        self.calibrate = list(calibrate or [])
        self._side: Optional[torch.cuda.Stream] = None
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.out = None
        self.captures = 0

    def capture(self) -> "GraphedStep":
        cur = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream()
        side = self._side
        side.wait_stream(cur)
        switch = _rz.capacity_binning(self.policy)
        switch.enable()
        try:
            with torch.cuda.stream(side):
                for prep in self.calibrate:
                    prep()
                    self.fn()
                for _ in range(8):          # until the capacity path ran clean at the learnt capacity
                    for _ in range(self.warmup):
                        self.fn()
                    side.synchronize()
                    try:
                        self.policy.check()
                        break
                    except _rz.CapacityOverflow:
                        continue
                else:
                    raise _rz.CapacityOverflow("the binning capacity did not settle during warm-up")
            # capture on the stream the warm-up ran on: autograd's AccumulateGrad nodes remember the stream they
            # were created on, and a node that syncs with another stream (worst case the legacy default stream,
            # cudaErrorStreamCaptureImplicit) invalidates the capture. Such a node only exists while an autograd
            # graph from before the warm-up is still alive; if a capture fails, drop what can be dropped and retry.
            for attempt in range(3):
                self.graph = torch.cuda.CUDAGraph()
                try:
                    with torch.cuda.graph(self.graph, stream=side):
                        self.out = self.fn()
                    break
                except RuntimeError:
                    self.graph = self.out = None
                    torch.cuda.set_stream(cur)     # (a failed capture_end leaves torch's stream context un-exited)
                    if attempt == 2:
                        raise
                    gc.collect()
                    torch.cuda.synchronize()
                    with torch.cuda.stream(side):
                        self.fn()
                    side.synchronize()
            cur.wait_stream(side)
        finally:
            switch.disable()
        self.captures += 1
        return self

    def replay(self):
        """Launch the captured step on the current stream; returns the captured output tensor(s) (static)."""
        if self.graph is None:
            self.capture()
        self.graph.replay()
        return self.out

    def verify(self) -> bool:
        """Call after a synchronisation point. False = the last replays overflowed the captured binning
        capacity (their results are invalid): the step was re-captured with a larger one, replay again."""
        try:
            self.policy.check()
            return True
        except _rz.CapacityOverflow:
            self.graph = None
            self.capture()
            return False
