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


class MultiViewStep:
    """V views of one training step in flight at once (the small-scene half).

    At the reference's own scene sizes (5 k - 100 k curve-Gaussians) one captured step is a chain of ~45 kernels of a
    few microseconds each: a single replay keeps a handful of the 148 SMs busy and is bound by the latency of the
    chain, not by work. Here the SAME step is captured V times - each copy with its own camera slot, target image,
    rasterizer state and gradient buffer - and the V graphs are replayed concurrently on V streams, so the chains of
    different views interleave on the GPU. Every copy computes exactly what a single `GraphedStep` computes for its
    view (same kernels, same launch geometry: per-view results are bit-identical); the copies' curve-parameter
    gradients are rows of one (V, n) buffer and are summed into the parameters' `.grad` after the replays, which is
    the multi-view accumulation of `parallel.FlatGrad` with the views running side by side instead of one after the
    other.

    body(cam, gt) -> loss runs one view: it must call `model.prepare_scaling_rot()`, `render(cam, ...)`, a loss and
    `.backward()`, reading the view from `cam` (a StaticCamera) and `gt` (a fixed tensor).
    """

    def __init__(self, params: Sequence[torch.Tensor], body: Callable[[StaticCamera, torch.Tensor], torch.Tensor],
                 like_cam, like_gt: torch.Tensor, views: int):
        self.params = list(params)
        self.V = int(views)
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flats = torch.zeros(self.V, n, dtype=self.params[0].dtype, device=dev)
        self.total = torch.zeros(n, dtype=self.params[0].dtype, device=dev)
        self.scams = [StaticCamera(like_cam) for _ in range(self.V)]
        self.gts = [torch.empty_like(like_gt) for _ in range(self.V)]
        self.streams = [torch.cuda.Stream(dev) for _ in range(self.V)]
        self.steps = []
        for k in range(self.V):
            def fn(k=k):
                self.flats[k].zero_()
                return body(self.scams[k], self.gts[k])
            self.steps.append(GraphedStep(fn))
        # all copies warm up and capture on ONE side stream: autograd remembers, per parameter, the stream its
        # AccumulateGrad node was created on and makes every backward wait for it - from a capture on any other stream
        # that is a dependency on uncaptured work (cudaErrorStreamCaptureIsolation). A graph does not remember the
        # stream it was captured on, so the copies still replay on V different streams.
        side = torch.cuda.Stream(dev)
        for step in self.steps:
            step._side = side

    def _bind(self, flat: torch.Tensor) -> None:
        off = 0
        for p in self.params:
            p.grad = flat[off:off + p.numel()].view_as(p)
            p._cg_direct_grad = True      # the library's backward kernels add straight into the bound buffer
            off += p.numel()

    def capture(self, calibrate_cams: Sequence = ()) -> "MultiViewStep":
        """Capture the V copies, one after the other; `calibrate_cams`: views that show the binning policy of every
        copy the largest tile-instance count it will meet."""
        for k, step in enumerate(self.steps):
            self._bind(self.flats[k])          # this copy's backward is captured writing into row k
            step.calibrate = [(lambda c=c, k=k: self.scams[k].load(c)) for c in calibrate_cams]
            step.capture()
        self._bind(self.total)                 # what the optimizer sees
        return self

    def replay(self, cams: Sequence, gts: Sequence[torch.Tensor]):
        """Run len(cams) <= V views concurrently; returns their (static) loss tensors. Afterwards the parameters'
        `.grad` (views of `self.total`) hold the SUM of the views' gradients."""
        nv = len(cams)
        if nv > self.V or nv != len(gts):
            raise ValueError(f"{nv} views for {self.V} captured copies")
        cur = torch.cuda.current_stream()
        for k in range(nv):
            self.scams[k].load(cams[k])
            self.gts[k].copy_(gts[k], non_blocking=True)
        outs = []
        for k in range(nv):
            s = self.streams[k]
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                outs.append(self.steps[k].replay())
        for k in range(nv):
            cur.wait_stream(self.streams[k])
        torch.sum(self.flats[:nv], dim=0, out=self.total)
        return outs

    def verify(self) -> bool:
        """After a synchronisation: False if a copy overflowed its captured binning capacity (it was re-captured)."""
        ok = True
        for k, step in enumerate(self.steps):
            try:
                step.policy.check()
            except _rz.CapacityOverflow:
                self._bind(self.flats[k])
                step.graph = None
                step.capture()
                self._bind(self.total)
                ok = False
        return ok
