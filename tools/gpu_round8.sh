#!/bin/bash
TAG=${1:-r01i}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout=600 > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -12 $OUT/pytest_$TAG.log
echo "== bench cfg2"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "exit $?"; python - <<PY
import json
d=json.load(open("$OUT/bench_$TAG.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["step_tensor_frac"])
print({k:round(v["ms_per_step"],2) for k,v in d["kernel_breakdown"].items()})
PY
tail -3 $OUT/bench_$TAG.err
echo "== bench cfg2 no overlap"; timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-overlap > $OUT/bench_noov_$TAG.json 2> $OUT/bench_noov_$TAG.err; python - <<PY
import json
d=json.load(open("$OUT/bench_noov_$TAG.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["achieved"])
print({k:round(v["ms_per_step"],2) for k,v in d["kernel_breakdown"].items()})
PY
echo "== kernels"; timeout 900 python tools/bench_kernels.py > $OUT/kernels_$TAG.json 2> $OUT/kernels_$TAG.err; echo "exit $?"; tail -5 $OUT/kernels_$TAG.err; python - <<PY
import json
d=json.load(open("$OUT/kernels_$TAG.json"))
for k in d["kernels"]:
    print("%-60s %-34s %9.3f ms %8.1f GB/s %5.1f%%" % (k["kernel"][:60], k["size"][:34], k["ms_median"], k["gbs"], 100*k["frac_of_hbm_peak"]))
PY
echo "== ncu hier_head / pool / topk / candidates"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"hier_head|pair_relu_pool_tiled|topk_match|candidates_kernel" -c 6 -o $OUT/prof_small_$TAG -f \
    python tools/bench_kernels.py --iters 1 --scale-images 512 > $OUT/ncu_small_$TAG.log 2>&1
echo "small capture exit $?"
