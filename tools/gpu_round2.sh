#!/bin/bash
TAG=${1:-r01b}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout=600 > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -15 $OUT/pytest_$TAG.log
for cfg in "" "--no-overlap" "--conv3-m-sub 1" "--chunk-pairs 32768"; do
  name=$(echo "bench_${TAG}${cfg}" | tr -d ' -')
  echo "== bench $cfg"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $cfg > $OUT/$name.json 2> $OUT/$name.err; echo "exit $?"
  python - <<PY
import json
try:
    d=json.load(open("$OUT/$name.json"))
    print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d["roofline"]["achieved"], {k:round(v["ms_per_step"],2) for k,v in d["kernel_breakdown"].items()}, d["clocks"])
except Exception as e:
    print("bad json", e); print(open("$OUT/$name.err").read()[-1500:])
PY
done
