#!/usr/bin/env bash
# Dev tool: GPU parity tests, then the bench line with the ring backward on (default) and off.
TAG=${1:-ring_ab}
OUT=gpurun_out/$TAG
mkdir -p $OUT
if [ "${2:-}" != "skip-tests" ]; then
  timeout -k 10 500 python -m pytest tests -m gpu -x -q -o faulthandler_timeout=120 > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log
  tail -15 $OUT/pytest.log
fi
for v in 1 0; do
  CURVEGS_BWD_RING=$v timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-small-scene --no-reference-step > $OUT/bench_ring$v.json 2> $OUT/bench_ring$v.err
  python - "$OUT/bench_ring$v.json" "$v" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("ring=" + sys.argv[2], d["value"], d["ms_per_step"], d["e2e"]["value"], d["stage_ms"])
except Exception as e:
    print("ring=" + sys.argv[2], "FAILED", e)
PY
done
