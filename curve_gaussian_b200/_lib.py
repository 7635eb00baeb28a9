"""ctypes binding of libcurvegs.so (the C ABI declared in include/curvegs.h).

There is no fallback: if the library is missing or a call fails, an exception
is raised. PyTorch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# CURVEGS_LIB selects another build of the same library (A/B timing of kernel variants, see build.build_variant)
LIB_PATH = os.environ.get("CURVEGS_LIB") or os.path.join(_HERE, "libcurvegs.so")

ABI_VERSION = 6


class RasterSettings(C.Structure):
    """struct cg_raster_settings"""

    _fields_ = [
        ("image_height", C.c_int32),
        ("image_width", C.c_int32),
        ("tanfovx", C.c_float),
        ("tanfovy", C.c_float),
        ("scale_modifier", C.c_float),
        ("render_geo", C.c_int32),
        ("debug", C.c_int32),
        ("antialiasing", C.c_int32),
        ("bg", C.c_void_p),
        ("viewmatrix", C.c_void_p),
        ("projmatrix", C.c_void_p),
    ]


class CurveGSError(RuntimeError):
    pass


_vp, _i64, _i32, _f32, _sz = C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_size_t
_SP = C.POINTER(RasterSettings)

# name -> (restype, argtypes); this table is also what tests check against the header.
SIGNATURES = {
    "cg_abi_version": (C.c_int, []),
    "cg_last_error": (C.c_char_p, []),
    "cg_launch_count": (C.c_uint64, []),
    "cg_profile_enable": (None, [C.c_int]),
    "cg_profile_reset": (None, []),
    "cg_profile_stage_count": (C.c_int, []),
    "cg_profile_stage_name": (C.c_char_p, [C.c_int]),
    "cg_profile_read": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "cg_raster_geom_bytes": (_sz, [_i64]),
    "cg_raster_img_bytes": (_sz, [_i32, _i32]),
    "cg_raster_bin_keep_bytes": (_sz, [_i64]),
    "cg_raster_bin_scratch_bytes": (_sz, [_i64, _i64]),
    "cg_raster_bwd_scratch_bytes": (_sz, [_i64]),
    "cg_raster_fwd_geom": (C.c_int, [_SP, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, C.POINTER(_i64), _vp]),
    "cg_raster_fwd_blend": (C.c_int, [_SP, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cg_raster_fwd_capacity": (C.c_int, [_SP, _i64, _i64] + [_vp] * 9 + [_sz] + [_vp] * 8),
    "cg_set_pdl": (None, [C.c_int]),
    "cg_raster_bwd": (C.c_int, [_SP, _i64, _i64] + [_vp] * 22),
    "cg_mark_visible": (C.c_int, [_i64, _vp, _vp, _vp, _vp, _vp]),
    "cg_raster_debug_fetch": (C.c_int, [C.c_int, _i64, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cg_sample_scratch_bytes": (_sz, [_i64, _i32]),
    "cg_sample_fwd": (C.c_int, [_i64, _i32, _vp, _vp, _vp, _vp, _f32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cg_sample_bwd": (C.c_int, [_i64, _i32, _vp, _vp, _vp, _vp, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp]),
    "cg_activate_fwd": (C.c_int, [_i64, _i32, _vp, _vp, _vp, _vp, _vp, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cg_activate_bwd": (C.c_int, [_i64, _i32, _vp, _vp, _vp, _vp, _vp, _f32, _vp, _vp] + [_vp] * 8 + [_i32, _vp]),
    "cg_ssim_fwd": (C.c_int, [_i32, _i32, _i32, _i32, _f32, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cg_ssim_bwd": (C.c_int, [_i32, _i32, _i32, _i32, _f32, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cg_edge_ssim_loss_stats_bytes": (_sz, []),
    "cg_edge_ssim_loss_fwd": (C.c_int, [_i32, _i32, _vp, _vp, _f32, _f32, _f32, _f32, _f32, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cg_edge_ssim_loss_bwd": (C.c_int, [_i32, _i32, _vp, _vp, _f32, _f32, _f32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cg_rotate_channels": (C.c_int, [_i64, _vp, _vp, _i32, _i32, _vp, _vp]),
    "cg_curve_smooth_scratch_bytes": (_sz, []),
    "cg_curve_smooth_fwd": (C.c_int, [_i64, _i32, _vp, _vp, _vp, _vp]),
    "cg_curve_smooth_bwd": (C.c_int, [_i64, _i32, _vp, _vp, _vp, _vp]),
    "cg_endpoint_conn_scratch_bytes": (_sz, [_i64]),
    "cg_endpoint_conn_fwd": (C.c_int, [_i64, _vp, _f32, _vp, _vp, _vp, _vp]),
    "cg_endpoint_conn_bwd": (C.c_int, [_i64, _vp, _vp, _vp, _vp, _vp]),
    "cg_knn_scratch_bytes": (_sz, [_i64]),
    "cg_knn_mean_dist2": (C.c_int, [_i64, _vp, _vp, _vp, _vp]),
}

_lib = None


def load() -> C.CDLL:
    """Load libcurvegs.so; raise if it is absent (no CPU/eager fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CurveGSError(
            f"{LIB_PATH} is missing: build it with `python -m curve_gaussian_b200.build` "
            "(or __graft_entry__.build()). There is no fallback path."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        if not hasattr(lib, name):
            raise CurveGSError(f"{LIB_PATH} does not export {name}; rebuild it")
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.cg_abi_version() != ABI_VERSION:
        raise CurveGSError(f"libcurvegs ABI {lib.cg_abi_version()} != expected {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def profile_read() -> dict:
    """{stage: (total_ms, calls)} accumulated since the last cg_profile_reset()."""
    lib = load()
    out = {}
    for i in range(lib.cg_profile_stage_count()):
        ms, n = C.c_double(0), C.c_uint64(0)
        lib.cg_profile_read(i, C.byref(ms), C.byref(n))
        if n.value:
            out[lib.cg_profile_stage_name(i).decode()] = (ms.value, int(n.value))
    return out


class _NullContext:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NULL = _NullContext()


def on_device(dev):
    """Context manager making `dev` the current CUDA device; free when it already is (the common case)."""
    import torch
    idx = dev.index
    if idx is None or idx == torch.cuda.current_device():
        return _NULL
    return torch.cuda.device(idx)


def stream(dev) -> int:
    """Raw cudaStream_t of torch's current stream on `dev` (what every cg_* call takes as `void* stream`)."""
    import torch
    idx = dev.index
    if idx is None:
        idx = torch.cuda.current_device()
    return torch._C._cuda_getCurrentRawStream(idx)


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().cg_last_error().decode("utf-8", "replace")
        raise CurveGSError(f"{what} failed (code {rc}): {msg}")


def ptr(t) -> int | None:
    """Device pointer of a torch tensor, or None (NULL) for absent/empty tensors."""
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()
