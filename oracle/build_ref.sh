#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY. Builds the UNMODIFIED reference CUDA extensions
# (diff-cur-rasterization, fused-ssim, simple-knn) for sm_100 from the sources
# where they lie under $REF, in a scratch copy under /tmp, and drops ONLY the
# resulting .so files into oracle/_ref/ (git-ignored, travels to the GPU box).
# The one patch applied to the scratch copy is `#include <cstdint>` in
# cuda_rasterizer/rasterizer_impl.h (gcc 13 no longer leaks uint32_t), which
# does not change any arithmetic. No reference source enters the repo.
set -euo pipefail
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
if [ ! -d "$REF/submodules" ]; then echo "no reference at $REF; keeping prebuilt $OUT"; exit 0; fi
mkdir -p "$OUT"
TMP=${REF_BUILD_TMP:-/tmp/curvegs_refbuild}
mkdir -p "$TMP"
for sub in diff-cur-rasterization fused-ssim simple-knn; do
  case $sub in
    diff-cur-rasterization) so=diff_cur_rasterization_C.so ;;
    fused-ssim) so=fused_ssim_cuda.so ;;
    simple-knn) so=simple_knn_C.so ;;
  esac
  if [ -f "$OUT/$so" ] && [ -z "${FORCE:-}" ]; then echo "$so present"; continue; fi
  rm -rf "$TMP/$sub"; cp -r "$REF/submodules/$sub" "$TMP/$sub"
  rm -rf "$TMP/$sub/build"
  if [ $sub = diff-cur-rasterization ]; then
    sed -i '0,/#include <iostream>/s//#include <cstdint>\n#include <iostream>/' "$TMP/$sub/cuda_rasterizer/rasterizer_impl.h"
  fi
  if [ $sub = simple-knn ]; then mkdir -p "$TMP/$sub/simple_knn"; fi
  (cd "$TMP/$sub" && TORCH_CUDA_ARCH_LIST=10.0 MAX_JOBS=${MAX_JOBS:-4} python setup.py build_ext > build.log 2>&1) || { tail -30 "$TMP/$sub/build.log"; exit 1; }
  built=$(find "$TMP/$sub/build" -name '*.so' | head -1)
  cp "$built" "$OUT/$so"
  echo "built $so"
done

# The reference's pure-Python side of the hot path (model sampling, render(), losses, the two extension wrappers),
# staged UNMODIFIED into the git-ignored oracle/_ref/py/ so that oracle/ref_step.py can drive the reference's real
# train.py step on the GPU box (where /root/reference does not exist). Nothing here enters the repository history.
PY=$OUT/py
mkdir -p "$PY/scene" "$PY/utils" "$PY/gaussian_renderer" "$PY/diff_cur_rasterization" "$PY/fused_ssim"
cp "$REF/scene/gaussian_curve_model.py" "$REF/scene/gaussian_model.py" "$PY/scene/"
cp "$REF/utils/general_utils.py" "$REF/utils/graphics_utils.py" "$REF/utils/sh_utils.py" "$REF/utils/system_utils.py" \
   "$REF/utils/loss_utils.py" "$PY/utils/"
cp "$REF/gaussian_renderer/__init__.py" "$PY/gaussian_renderer/__init__.py"
cp "$REF/submodules/diff-cur-rasterization/diff_cur_rasterization/__init__.py" "$PY/diff_cur_rasterization/__init__.py"
cp "$REF/submodules/fused-ssim/fused_ssim/__init__.py" "$PY/fused_ssim/__init__.py"
echo "staged reference python under $PY"
