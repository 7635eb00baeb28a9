"""Drop-in for `simple_knn._C.distCUDA2` (submodules/simple-knn/spatial.cu:15-26)."""
from __future__ import annotations

import torch

from . import _lib


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    if not points.is_cuda:
        raise _lib.CurveGSError("distCUDA2 needs a CUDA tensor; there is no CPU path")
    pts = points.float().contiguous()
    P = pts.shape[0]
    out = torch.zeros((P,), dtype=torch.float32, device=pts.device)
    if P == 0:
        return out
    scratch = torch.empty(lib.cg_knn_scratch_bytes(P), dtype=torch.uint8, device=pts.device)
    with _lib.on_device(pts.device):
        _lib.check(lib.cg_knn_mean_dist2(P, pts.data_ptr(), out.data_ptr(), scratch.data_ptr(),
                                         _lib.stream(pts.device)), "cg_knn_mean_dist2")
    return out
