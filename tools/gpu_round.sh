#!/bin/bash
# one GPU call: sparse parity tests (fail fast), the whole GPU suite, A/B of the conv3_1 modes, the default bench line,
# the ncu launch list of the default command
TAG=${1:-r01v}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv,noheader
echo "== pytest sparse"; timeout 600 python -m pytest tests/test_gpu_sparse.py -q -x --timeout=300 > $OUT/pytest_sparse_$TAG.log 2>&1; rc=$?; echo "pytest exit $rc"; tail -30 $OUT/pytest_sparse_$TAG.log
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q --timeout=600 > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?"; tail -15 $OUT/pytest_gpu_$TAG.log
for mode in ${MODES:-shared4 shared8 blocks4 dense}; do
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
echo "== bench default"; timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "exit $?"; head -c 3000 $OUT/bench_$TAG.json; echo; tail -3 $OUT/bench_$TAG.err
if [ -z "$NO_NCU" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_bench_$TAG.log 2>&1
echo "launch list exit $?"; wc -l $OUT/launches_$TAG.csv
fi
