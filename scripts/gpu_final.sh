#!/usr/bin/env bash
# Round-end evidence run: parity tests, smoke, bench (ours + reference arm), ncu launch list and one
# ncu --set full pass over every libcurvegs kernel of one step.
set -u
TAG=${1:-final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
# SASS evidence: TMA bulk copies (UBLKCP + SYNCS mbarrier), async gathers (LDGSTS), vector reductions (REDG ... F32x4)
cuobjdump -sass curve_gaussian_b200/libcurvegs.so | grep -E "Function :|UBLKCP|SYNCS\.|LDGSTS|REDG.*F32x4|FFMA2" | awk '/Function/{f=$0; next} {m=$2; if (m ~ /^@/) m=$3; c[f" | "m]++} END{for(k in c) print c[k], k}' | sort -k3 > $OUT/sass_excerpt.txt 2>&1
[ -x scripts/micro/cub_sort_yardstick ] && scripts/micro/cub_sort_yardstick > $OUT/cub_sort_yardstick.txt 2>&1
timeout -k 10 400 python -m pytest tests -m gpu -x -q -o faulthandler_timeout=120 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log
if ! grep -q "pytest exit 0" $OUT/pytest_gpu.log; then echo "GPU tests failed: stopping"; exit 1; fi
timeout -k 10 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -2 $OUT/smoke.log
timeout -k 10 400 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
timeout -k 10 400 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref arm exit $?"
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 2 --warmup 3 --views 2 --no-cpu-baseline --no-small-scene --no-reference-step > $OUT/ncu_launch_bench.log 2>&1
# one --set full pass over the kernels that make up >90 % of the step (FULL=1: every libcurvegs kernel of one step)
if [ "${FULL:-0}" = "1" ]; then
  KRE='^(activate_|blend_|bin_|emit_keys|tile_ranges|perm_block_sums|preprocess_fwd|preprocess_bwd|sample_|scan_block_sums|sort_|ssim_)'; SKIP=200; CNT=42
else
  KRE='^(blend_|bin_fill|bin_count|sort_onesweep|preprocess_fwd|preprocess_bwd|sample_bwd_point|ssim_fwd|emit_keys)'; SKIP=60; CNT=14
fi
timeout -k 10 500 ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s $SKIP -c $CNT -o $OUT/full \
  python bench.py --steps 2 --warmup 3 --views 2 --no-cpu-baseline --no-small-scene --no-reference-step > $OUT/ncu_full_bench.log 2>&1
ls -la $OUT
# A/B of kernel variants built by `python -m curve_gaussian_b200.build --variant ...` (dev only; last, so that
# running out of time here costs nothing above)
for f in curve_gaussian_b200/variants/libcurvegs_*.so; do
  [ -e "$f" ] || continue
  n=$(basename $f .so); n=${n#libcurvegs_}
  CURVEGS_LIB=$PWD/$f timeout -k 10 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-small-scene --no-reference-step > $OUT/bench_variant_$n.json 2> $OUT/bench_variant_$n.err
  python - "$OUT/bench_variant_$n.json" "$n" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("variant", sys.argv[2], d["value"], d["ms_per_step"], d["e2e"]["value"], d["stage_ms"])
except Exception as e:
    print("variant", sys.argv[2], "FAILED", e)
PY
done
