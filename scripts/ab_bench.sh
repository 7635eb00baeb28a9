#!/usr/bin/env bash
# Dev tool: GPU parity tests once, then the bench line (no CPU baseline) for the default build and for every
# variants/libcurvegs_*.so; prints value / ms / e2e / stage table per build.
# usage: scripts/ab_bench.sh <tag> [skip-tests]
set -u
TAG=${1:-ab}
OUT=gpurun_out/$TAG
mkdir -p $OUT
if [ "${2:-}" != "skip-tests" ]; then
  timeout -k 10 300 python -m pytest tests -m gpu -x -q -o faulthandler_timeout=60 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
  tail -3 $OUT/pytest_gpu.log
fi
run() {
  local name=$1; shift
  timeout -k 10 200 env "$@" python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-small-scene --no-reference-step > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  python - "$OUT/bench_$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], d["value"], d["ms_per_step"], d["e2e"]["value"], d["stage_ms"])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
run default CURVEGS_LIB=
for f in curve_gaussian_b200/variants/libcurvegs_*.so; do
  [ -e "$f" ] || continue
  n=$(basename $f .so); n=${n#libcurvegs_}
  run $n CURVEGS_LIB=$PWD/$f
done
