#!/usr/bin/env bash
# Dev tool: GPU parity tests, then the bench line with super-tile binning (default) and with the sort path.
TAG=${1:-bin_ab}
OUT=gpurun_out/$TAG
mkdir -p $OUT
if [ "${2:-}" != "skip-tests" ]; then
  timeout -k 10 500 python -m pytest tests -m gpu -x -q -o faulthandler_timeout=120 > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log
  tail -15 $OUT/pytest.log
fi
for v in supertile sort; do
  CURVEGS_BINNING=$v timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-small-scene --no-reference-step > $OUT/bench_$v.json 2> $OUT/bench_$v.err
  python - "$OUT/bench_$v.json" "$v" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], d["value"], d["ms_per_step"], d["e2e"]["value"], d["stage_ms"])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
done
