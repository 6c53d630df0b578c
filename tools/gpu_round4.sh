#!/bin/bash
# pytest (all GPU parity tests) + cfg3 bench + per-kernel HBM microbench
TAG=${1:-r01d}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout=600 > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -30 $OUT/pytest_$TAG.log
echo "== bench cfg3"; timeout 900 python bench.py --workload cfg3 --steps 10 --warmup 3 > $OUT/bench_cfg3_$TAG.json 2> $OUT/bench_cfg3_$TAG.err; echo "exit $?"; head -c 2500 $OUT/bench_cfg3_$TAG.json; echo; tail -3 $OUT/bench_cfg3_$TAG.err
echo "== kernels"; timeout 900 python tools/bench_kernels.py > $OUT/kernels_$TAG.json 2> $OUT/kernels_$TAG.err; echo "exit $?"; tail -5 $OUT/kernels_$TAG.err; python - <<PY
import json
d=json.load(open("$OUT/kernels_$TAG.json"))
for k in d["kernels"]:
    print("%-60s %-34s %9.3f ms %8.1f GB/s %5.1f%%" % (k["kernel"][:60], k["size"][:34], k["ms_median"], k["gbs"], 100*k["frac_of_hbm_peak"]))
PY
