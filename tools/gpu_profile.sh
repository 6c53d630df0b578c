#!/bin/bash
# ncu evidence for one round: (1) per-launch device time of every kernel of a short bench run,
# (2) one --set full capture of the dominant kernels (conv3_1, fc1, fc2 of the first pair chunk).
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_bench_$TAG.log 2>&1
echo "launch list exit $?"; wc -l $OUT/launches_$TAG.csv
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 3 -c 3 -o $OUT/prof_$TAG -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_full_$TAG.log 2>&1
echo "full capture exit $?"; ls -la $OUT/prof_$TAG.ncu-rep
timeout 900 ncu --set full --clock-control none -k regex:"pair_relu_pool|topk_match|hier_head|box_select|candidates" -c 6 -o $OUT/prof_small_$TAG -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_small_$TAG.log 2>&1
echo "small capture exit $?"
