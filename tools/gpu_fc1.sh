#!/bin/bash
# shared-footprint fc1: parity tests first (fail fast), then A/B against the dense fc1 on the same box
TAG=${1:-r01w}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest fc1_shared + sparse"; timeout 900 python -m pytest tests/test_gpu_fc1_shared.py tests/test_gpu_sparse.py -q --timeout=300 > $OUT/pytest_fc1_$TAG.log 2>&1; rc=$?; echo "pytest exit $rc"; tail -40 $OUT/pytest_fc1_$TAG.log
for mode in ${MODES:-shared dense}; do
  echo "== bench fc1 $mode"; timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --fc1 $mode > $OUT/bench_fc1${mode}_$TAG.json 2> $OUT/bench_fc1${mode}_$TAG.err; echo "exit $?"
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_fc1${mode}_$TAG.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "conv3_blocks_per_step", "fc1_cells_per_step")}, d["e2e"]["value"], d["clocks"], {k: round(v["ms_per_step"], 2) for k, v in d["kernel_breakdown"].items()}, d["roofline"]["kernel"][:40], d["roofline"]["achieved"], d["roofline"].get("executed_fraction"), d["recall"])
except Exception as e:
    print("no line:", e); print(open("$OUT/bench_fc1${mode}_$TAG.err").read()[-2500:])
PY
done
