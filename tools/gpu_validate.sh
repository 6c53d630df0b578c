#!/bin/bash
# round-end validation exactly as the driver does it: all GPU parity tests, the default bench line, the reference arm
TAG=${1:-r01}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout=600 > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -40 $OUT/pytest_$TAG.log
echo "== bench"; timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "exit $?"; head -c 1500 $OUT/bench_$TAG.json; echo; tail -3 $OUT/bench_$TAG.err
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "exit $?"; head -c 900 $OUT/bench_ref_$TAG.json
