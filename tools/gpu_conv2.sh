#!/bin/bash
# conv2_1 halves on the box footprint: kernel-level parity (fail fast), then A/B on the same box
TAG=${1:-r01N}
OUT=gpurun_out; mkdir -p $OUT
timeout 100 python -m pytest tests/test_gpu_sparse.py -q -x -k conv2_halves --timeout=90 > $OUT/pytest_conv2_$TAG.log 2>&1; rc=$?; echo "pytest exit $rc"; tail -6 $OUT/pytest_conv2_$TAG.log
if [ $rc -ne 0 ]; then exit $rc; fi
for mode in 1 0; do
  HC_CONV2_SPARSE=$mode timeout 60 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $OUT/bench_conv2s${mode}_$TAG.json 2> $OUT/bench_conv2s${mode}_$TAG.err; echo "exit $?"
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_conv2s${mode}_$TAG.json"))
    print($mode, {k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["clocks"]["sm_mhz"], {k: round(v["ms_per_step"], 2) for k, v in d["kernel_breakdown"].items()}, d["recall"])
except Exception as e:
    print("no line:", e); print(open("$OUT/bench_conv2s${mode}_$TAG.err").read()[-1200:])
PY
done
