#!/usr/bin/env bash
# Round-end evidence run: parity tests, smoke, bench (ours + reference arm), ncu launch list and one
# ncu --set full pass over every libcurvegs kernel of one step.
set -u
TAG=${1:-final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
timeout -k 10 400 python -m pytest tests -m gpu -x -q -o faulthandler_timeout=120 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log
if ! grep -q "pytest exit 0" $OUT/pytest_gpu.log; then echo "GPU tests failed: stopping"; exit 1; fi
timeout -k 10 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -2 $OUT/smoke.log
timeout -k 10 300 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
timeout -k 10 400 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref arm exit $?"
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 2 --warmup 3 --views 2 --no-cpu-baseline > $OUT/ncu_launch_bench.log 2>&1
timeout -k 10 700 ncu --set full --clock-control none --import-source on -k regex:"^(activate_|blend_|emit_keys|gather_records|init_depth_keys|perm_block_sums|preprocess_fwd|preprocess_bwd|sample_|scan_block_sums|sort_|ssim_)" -s 200 -c 40 -o $OUT/full \
  python bench.py --steps 2 --warmup 3 --views 2 --no-cpu-baseline > $OUT/ncu_full_bench.log 2>&1
ls -la $OUT
