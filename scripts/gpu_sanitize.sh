#!/usr/bin/env bash
# compute-sanitizer memcheck / synccheck / racecheck over the small GPU parity tests (run through gpurun).
for tool in memcheck synccheck; do
  echo "=== $tool"
  timeout -k 10 400 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_raster_golden.py \
    tests/test_gpu_render_pipeline.py tests/test_gpu_regularizers.py tests/test_gpu_loss.py -x -q -m gpu \
    -k "not 1080 and not 6000 and not 2000" 2>&1 | tail -3
done
echo "=== racecheck"
timeout -k 10 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_raster_golden.py -x -q -m gpu 2>&1 | tail -3
echo "=== memcheck: capacity binning + CUDA-graph capture / replay (round 1: the capture failed under the sanitizer)"
timeout -k 10 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_capacity_graph.py -x -q -m gpu 2>&1 | tail -5
echo "=== memcheck: binning edge cases against the reference rasterizer (297 super-tiles, unstaged fill, ties / screen-filling splats)"
timeout -k 10 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_raster_vs_reference.py -x -q -m gpu \
  -k "(big_splats or wide or ties) and forward_backward" 2>&1 | tail -4
