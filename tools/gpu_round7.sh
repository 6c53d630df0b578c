#!/bin/bash
TAG=${1:-r01h}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -x > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -8 $OUT/pytest_$TAG.log
echo "== kernels"; timeout 900 python tools/bench_kernels.py > $OUT/kernels_$TAG.json 2> $OUT/kernels_$TAG.err; echo "exit $?"; tail -5 $OUT/kernels_$TAG.err; python - <<PY
import json
d=json.load(open("$OUT/kernels_$TAG.json"))
for k in d["kernels"]:
    print("%-60s %-34s %9.3f ms %8.1f GB/s %5.1f%%" % (k["kernel"][:60], k["size"][:34], k["ms_median"], k["gbs"], 100*k["frac_of_hbm_peak"]))
PY
bash tools/gpu_profile2.sh $TAG
