"""Loads the UNMODIFIED reference CUDA extensions built by oracle/build_ref.sh
into oracle/_ref/ (test infrastructure only; never imported by the product)."""
import importlib.machinery
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def _load(modname, filename):
    path = os.path.join(REF_DIR, filename)
    if not os.path.exists(path):
        return None
    import torch  # noqa: F401  (the extensions link against libtorch)
    loader = importlib.machinery.ExtensionFileLoader(modname, path)
    spec = importlib.util.spec_from_loader(modname, loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    return mod


_cache = {}


def ref_rasterizer():
    if "r" not in _cache:
        # both extensions export PyInit__C; distinct dotted names keep CPython's extension cache from aliasing them
        _cache["r"] = _load("ref_diff_cur_rasterization._C", "diff_cur_rasterization_C.so")
    return _cache["r"]


def ref_ssim():
    if "s" not in _cache:
        _cache["s"] = _load("fused_ssim_cuda", "fused_ssim_cuda.so")
    return _cache["s"]


def ref_knn():
    if "k" not in _cache:
        _cache["k"] = _load("ref_simple_knn._C", "simple_knn_C.so")
    return _cache["k"]
