#!/bin/bash
# conv patch pipeline: dense parity first (short timeout: a barrier bug would hang), then A/B timing against the per-tap pipeline
TAG=${1:-r01k}
OUT=gpurun_out; mkdir -p $OUT
echo "== dense parity (patch pipeline)"; timeout 240 python -m pytest tests/test_gpu_dense.py -m gpu -q -x --timeout=120 > $OUT/pytest_dense_$TAG.log 2>&1; rc=$?; echo "exit $rc"; tail -8 $OUT/pytest_dense_$TAG.log
if [ $rc -ne 0 ]; then echo "dense parity failed - stopping"; exit 1; fi
echo "== probe patch"; PROBE_PAIRS=4096 timeout 300 python tools/gpu_probe_conv3.py 2>&1 | grep -v "^\[{" | tail -16
echo "== probe taps"; HC_CONV_TAPS=1 PROBE_PAIRS=4096 timeout 300 python tools/gpu_probe_conv3.py 2>&1 | grep -v "^\[{" | grep "m_sub=2 epi=pool\|m_sub=1 epi=bf16"
echo "== full pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout=600 > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -6 $OUT/pytest_$TAG.log
for mode in patch taps; do
  echo "== bench cfg2 $mode"; if [ $mode = taps ]; then export HC_CONV_TAPS=1; else unset HC_CONV_TAPS; fi
  timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_${mode}_$TAG.json 2> $OUT/bench_${mode}_$TAG.err; echo "exit $?"; python - <<PY
import json
d=json.load(open("$OUT/bench_${mode}_$TAG.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["step_tensor_frac"], d["clocks"]["sm_mhz"])
print({k:round(v["ms_per_step"],2) for k,v in d["kernel_breakdown"].items()})
PY
done
unset HC_CONV_TAPS
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
