#!/bin/bash
# CTA-pair weight multicast in block mode (HC_CONV3_PAIRS=1): parity first under a hard timeout (a protocol bug would hang), then A/B
TAG=${1:-r01E}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest sparse + fc1_shared with CTA pairs"
HC_CONV3_PAIRS=1 timeout -s KILL 420 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_fc1_shared.py -q -x --timeout=300 > $OUT/pytest_pairs_$TAG.log 2>&1; rc=$?; echo "pytest exit $rc"; tail -15 $OUT/pytest_pairs_$TAG.log
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
if [ $rc -ne 0 ]; then exit $rc; fi
for mode in 1 0 1 0; do
  echo "== bench HC_CONV3_PAIRS=$mode"
  HC_CONV3_PAIRS=$mode timeout -s KILL 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $OUT/bench_pairs${mode}_$TAG.json 2> $OUT/bench_pairs${mode}_$TAG.err; echo "exit $?"
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_pairs${mode}_$TAG.json"))
    print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["clocks"], {k: round(v["ms_per_step"], 2) for k, v in d["kernel_breakdown"].items()}, d["roofline"]["achieved"], d["recall"]["R@20/50/100"])
except Exception as e:
    print("no line:", e); print(open("$OUT/bench_pairs${mode}_$TAG.err").read()[-2500:])
PY
done
