#!/bin/bash
# block-sparse conv3_1: parity tests first (fail fast), then A/B of the three conv3 modes on the same box
TAG=${1:-r01u}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest sparse"; timeout 600 python -m pytest tests/test_gpu_sparse.py -q -x --timeout=300 > $OUT/pytest_sparse_$TAG.log 2>&1; rc=$?; echo "pytest exit $rc"; tail -30 $OUT/pytest_sparse_$TAG.log
if [ $rc -ne 0 ]; then exit $rc; fi
for mode in ${MODES:-blocks4 shared8 shared4}; do
  echo "== bench $mode"; timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --conv3 $mode > $OUT/bench_${mode}_$TAG.json 2> $OUT/bench_${mode}_$TAG.err; echo "exit $?"
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_${mode}_$TAG.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "conv3_blocks_per_step")}, d["e2e"]["value"], d["clocks"], {k: round(v["ms_per_step"], 2) for k, v in d["kernel_breakdown"].items()}, d["roofline"]["achieved"], d["roofline"].get("executed_fraction"), d["recall"])
except Exception as e:
    print("no line:", e); print(open("$OUT/bench_${mode}_$TAG.err").read()[-1500:])
PY
done
