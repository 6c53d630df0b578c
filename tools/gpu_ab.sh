#!/bin/bash
# A/B runs on one box: parity tests (fail fast) then bench lines for each "NAME|ENV|ARGS" entry of $RUNS (default set below)
TAG=${1:-r01z}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest fc1_shared + sparse"; timeout 900 python -m pytest tests/test_gpu_fc1_shared.py tests/test_gpu_sparse.py -q -x --timeout=300 > $OUT/pytest_fc1_$TAG.log 2>&1; rc=$?; echo "pytest exit $rc"; tail -30 $OUT/pytest_fc1_$TAG.log
if [ $rc -ne 0 ]; then exit $rc; fi
IFS=';' read -ra LIST <<< "${RUNS:-g9_44|HC_FC1_GROUP_M=9|--conv3 shared44;g18_44|HC_FC1_GROUP_M=18|--conv3 shared44;g37_44|HC_FC1_GROUP_M=37|--conv3 shared44;g9_4|HC_FC1_GROUP_M=9|--conv3 shared4}"
for item in "${LIST[@]}"; do
  IFS='|' read -r name envs args <<< "$item"
  echo "== bench $name ($envs $args)"
  env $envs timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline $args > $OUT/bench_${name}_$TAG.json 2> $OUT/bench_${name}_$TAG.err; echo "exit $?"
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_${name}_$TAG.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "conv3_blocks_per_step", "fc1_cells_per_step")}, d["e2e"]["value"], d["clocks"], {k: round(v["ms_per_step"], 2) for k, v in d["kernel_breakdown"].items()}, d["roofline"]["kernel"][:40], d["roofline"]["achieved"], d["recall"]["R@20/50/100"])
except Exception as e:
    print("no line:", e); print(open("$OUT/bench_${name}_$TAG.err").read()[-2500:])
PY
done
