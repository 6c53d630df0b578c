#!/bin/bash
# all GPU parity tests + smoke + the default bench line (driver order), nothing else
TAG=${1:-r01}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout=600 > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?"; tail -8 $OUT/pytest_gpu_$TAG.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench"; timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "exit $?"; head -c 2600 $OUT/bench_$TAG.json; echo; tail -3 $OUT/bench_$TAG.err
