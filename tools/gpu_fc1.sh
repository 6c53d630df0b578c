#!/bin/bash
# shared-footprint fc1 + block shapes: parity tests first (fail fast), then A/B on the same box, then ncu evidence of the default
TAG=${1:-r01x}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest fc1_shared + sparse"; timeout 900 python -m pytest tests/test_gpu_fc1_shared.py tests/test_gpu_sparse.py -q --timeout=300 > $OUT/pytest_fc1_$TAG.log 2>&1; rc=$?; echo "pytest exit $rc"; tail -40 $OUT/pytest_fc1_$TAG.log
for mode in ${MODES:-shared44 shared4}; do
  echo "== bench conv3 $mode"; timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --conv3 $mode > $OUT/bench_${mode}_$TAG.json 2> $OUT/bench_${mode}_$TAG.err; echo "exit $?"
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_${mode}_$TAG.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "conv3_blocks_per_step", "fc1_cells_per_step")}, d["e2e"]["value"], d["clocks"], {k: round(v["ms_per_step"], 2) for k, v in d["kernel_breakdown"].items()}, d["roofline"]["kernel"][:40], d["roofline"]["achieved"], d["roofline"].get("executed_fraction"), d["recall"])
except Exception as e:
    print("no line:", e); print(open("$OUT/bench_${mode}_$TAG.err").read()[-2500:])
PY
done
if [ -z "$NO_NCU" ] && [ $rc -eq 0 ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_bench_$TAG.log 2>&1
echo "launch list exit $?"; wc -l $OUT/launches_$TAG.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -c 22 -o $OUT/prof_dense_$TAG -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_dense_$TAG.log 2>&1
echo "dense capture exit $?"; ls -la $OUT/*.ncu-rep
fi
