#!/bin/bash
# round-end validation as the driver does it: all GPU parity tests, smoke(), the default bench line, the reference arm, plus the ncu
# launch list of the default command and the cfg3 line
TAG=${1:-r01}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout=600 > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?"; tail -25 $OUT/pytest_gpu_$TAG.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench"; timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "exit $?"; head -c 1200 $OUT/bench_$TAG.json; echo; tail -3 $OUT/bench_$TAG.err
echo "== bench cfg3"; timeout 600 python bench.py --workload cfg3 --no-cpu-baseline > $OUT/bench_cfg3_$TAG.json 2> $OUT/bench_cfg3_$TAG.err; echo "exit $?"; head -c 600 $OUT/bench_cfg3_$TAG.json; echo; tail -3 $OUT/bench_cfg3_$TAG.err
echo "== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_bench_$TAG.log 2>&1; echo "exit $?"; wc -l $OUT/launches_$TAG.csv
if [ -z "$NO_REF" ]; then
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "exit $?"; head -c 900 $OUT/bench_ref_$TAG.json
fi
if [ -n "$KERNELS" ]; then
echo "== per-kernel bench"; timeout 600 python tools/bench_kernels.py > $OUT/kernels_$TAG.json 2> $OUT/kernels_$TAG.err; echo "exit $?"; tail -2 $OUT/kernels_$TAG.err
fi
