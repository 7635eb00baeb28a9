#!/usr/bin/env bash
# One GPU-box visit: parity tests, bench line, ncu launch list, ncu --set full of the hot kernels.
# usage: scripts/gpu_check.sh <tag> [kernel-regex]
set -u
TAG=${1:-run}
KREGEX=${2:-'blend_bwd|blend_fwd'}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
timeout -k 10 300 python -m pytest tests -m gpu -x -q -o faulthandler_timeout=60 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
if ! grep -q "pytest exit 0" $OUT/pytest_gpu.log; then echo "GPU tests failed: skipping bench and profiles"; exit 1; fi
timeout -k 10 300 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
cat $OUT/bench.json
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 2 --warmup 3 --views 2 --no-cpu-baseline > $OUT/ncu_launch_bench.log 2>&1
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:"$KREGEX" -s 8 -c 4 -o $OUT/full \
  python bench.py --steps 2 --warmup 3 --views 2 --no-cpu-baseline > $OUT/ncu_full_bench.log 2>&1
ls -la $OUT
