#!/bin/bash
# after kernel changes: parity first, then bench cfg2, kernel microbench, conv3 probe
TAG=${1:-r01e}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -x > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -15 $OUT/pytest_$TAG.log
echo "== bench cfg2"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "exit $?"; python - <<PY
import json
d=json.load(open("$OUT/bench_$TAG.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["step_tensor_frac"])
print({k:round(v["ms_per_step"],2) for k,v in d["kernel_breakdown"].items()})
PY
tail -3 $OUT/bench_$TAG.err
echo "== kernels"; timeout 900 python tools/bench_kernels.py > $OUT/kernels_$TAG.json 2> $OUT/kernels_$TAG.err; echo "exit $?"; tail -5 $OUT/kernels_$TAG.err; python - <<PY
import json
d=json.load(open("$OUT/kernels_$TAG.json"))
for k in d["kernels"]:
    print("%-60s %-34s %9.3f ms %8.1f GB/s %5.1f%%" % (k["kernel"][:60], k["size"][:34], k["ms_median"], k["gbs"], 100*k["frac_of_hbm_peak"]))
PY
echo "== probe"; PROBE_PAIRS=4096 timeout 300 python tools/gpu_probe_conv3.py 2>&1 | grep -v "^\[{" | tail -20
