#!/bin/bash
# ncu evidence: (1) launch list of a short cfg2 bench, (2) --set full of the dense kernels of the first pair chunks,
# (3) --set full of the HBM/latency-bound kernels, (4) the same for the SGB twin.  Summarise here (no GPU needed) with
#   python tools/ncu_summarize.py gpurun_out/prof_*_TAG.ncu-rep > profiles/ncu_summary_TAG.json
TAG=${1:-r01s}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_bench_$TAG.log 2>&1
echo "launch list exit $?"; wc -l $OUT/launches_$TAG.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 1 -c 8 -o $OUT/prof_dense_$TAG -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_dense_$TAG.log 2>&1
echo "dense capture exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pair_relu_pool_tiled|topk_match|hier_head|box_select|candidates_kernel|pairs_fill|box_label" -c 8 -o $OUT/prof_small_$TAG -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_small_$TAG.log 2>&1
echo "small capture exit $?"
STEPS=1 timeout 600 ncu --set full --clock-control none -k regex:"sgb_|tc_gemm" -s 14 -c 7 -o $OUT/prof_sgb_$TAG -f \
    python tools/bench_sgb.py > $OUT/ncu_sgb_$TAG.log 2>&1
echo "sgb capture exit $?"
ls -la $OUT/*.ncu-rep
