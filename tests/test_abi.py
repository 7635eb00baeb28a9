"""CPU: the C-ABI library loads and exports every symbol include/curvegs.h declares
(no compute calls), size queries behave, and the product never imports the oracle."""
import os
import re
import subprocess

from curve_gaussian_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "curvegs.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cg_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_header_symbols():
    build.build()
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 19
    for s in syms:
        assert hasattr(lib, s), s
    assert sorted(_lib.SIGNATURES) == syms
    assert lib.cg_abi_version() == _lib.ABI_VERSION


def test_dynamic_symbol_table_is_c_abi():
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    for s in header_symbols():
        assert s in exported, s


def test_size_queries():
    lib = _lib.load()
    assert lib.cg_raster_geom_bytes(0) >= 0
    a, b = lib.cg_raster_geom_bytes(1000), lib.cg_raster_geom_bytes(2000)
    assert 0 < a < b and a % 128 == 0
    assert lib.cg_raster_img_bytes(1920, 1080) >= 1920 * 1080 * 8
    assert lib.cg_raster_bin_keep_bytes(1000) >= 1000 * 52
    assert lib.cg_raster_bin_scratch_bytes(100, 1000) >= 1000 * 16
    assert lib.cg_raster_geom_bytes(1000) >= 1000 * (40 + 16)   # per-Gaussian state + the depth-sort buffers
    assert lib.cg_raster_bwd_scratch_bytes(10) == 320
    assert lib.cg_sample_scratch_bytes(10, 12) >= 32
    assert lib.cg_knn_scratch_bytes(3375) > 0


def test_bad_arguments_return_error_codes_not_crashes():
    lib = _lib.load()
    import ctypes as C
    s = _lib.RasterSettings()
    R = C.c_int64(0)
    rc = lib.cg_raster_fwd_geom(C.byref(s), 10, None, None, None, None, None, None, None, None, None, 0, C.byref(R), None)
    assert rc == -1
    assert b"image size" in lib.cg_last_error()


def test_capacity_forward_rejects_bad_arguments_before_touching_the_gpu():
    lib = _lib.load()
    import ctypes as C
    s = _lib.RasterSettings()
    s.image_width, s.image_height = 64, 48
    s.bg = s.viewmatrix = s.projmatrix = 0x1000          # never dereferenced: argument checks come first
    args = [None] * 9 + [0] + [None] * 8
    assert lib.cg_raster_fwd_capacity(C.byref(s), 0, 100, *args) == -1 and b"P" in lib.cg_last_error()
    assert lib.cg_raster_fwd_capacity(C.byref(s), 10, 0, *args) == -1 and b"R_cap" in lib.cg_last_error()
    assert lib.cg_raster_fwd_capacity(C.byref(s), 10, 1 << 30, *args) == -1
    assert lib.cg_raster_fwd_capacity(C.byref(s), 10, 100, *args) == -1 and b"means3D" in lib.cg_last_error()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "curve_gaussian_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("oracle/_ref", "").replace("oracle (", "") or f == "math.cuh", f
