#!/bin/bash
# One gpurun call: diagnostics -> GPU tests -> bench -> ncu launch list + full capture of the top kernel.
# Usage (under gpurun): bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > $OUT/gpu_$TAG.txt 2>&1
python -c "import torch; print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))" >> $OUT/gpu_$TAG.txt 2>&1
echo "== diag" ; timeout 600 python tools/gpu_diag.py > $OUT/diag_$TAG.log 2>&1; echo "diag exit $?"; tail -40 $OUT/diag_$TAG.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout=600 > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -30 $OUT/pytest_$TAG.log
echo "== pytest gpu (continue past failures)"; timeout 1500 python -m pytest tests -m gpu -q --timeout=600 > $OUT/pytest_all_$TAG.log 2>&1; tail -15 $OUT/pytest_all_$TAG.log
echo "== smoke"; timeout 600 python __graft_entry__.py smoke > $OUT/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -5 $OUT/smoke_$TAG.log
echo "== bench"; timeout 1200 python bench.py --steps 3 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"; cat $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
