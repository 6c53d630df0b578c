#!/bin/bash
# compute-sanitizer over the small-shape GPU tests: memcheck (global/shared OOB, misaligned) and racecheck (shared-memory hazards
# in the sort / scan / compaction kernels).  The dense tcgen05 tests are included in memcheck only at their smallest shapes.
TAG=${1:-r01}
OUT=gpurun_out; mkdir -p $OUT
SMALL="tests/test_gpu_frontend.py tests/test_gpu_variants.py tests/test_gpu_eval.py"
echo "== memcheck (integer + head kernels)"
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 66 --print-limit 20 python -m pytest $SMALL -m gpu -q -x --timeout=900 > $OUT/memcheck_$TAG.log 2>&1
echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|error" $OUT/memcheck_$TAG.log | tail -8
echo "== memcheck (dense tcgen05 kernels, small shapes)"
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 66 --print-limit 20 python -m pytest tests/test_gpu_dense.py -m gpu -q -x --timeout=900 -k "plain_gemm_f32 or implicit_conv_bf16 or pack_pixels or tiled_pair_pool" > $OUT/memcheck_dense_$TAG.log 2>&1
echo "memcheck dense exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|error" $OUT/memcheck_dense_$TAG.log | tail -8
echo "== racecheck (integer kernels)"
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 66 --print-limit 20 python -m pytest tests/test_gpu_frontend.py tests/test_gpu_eval.py -m gpu -q -x --timeout=900 -k "not model" > $OUT/racecheck_$TAG.log 2>&1
echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" $OUT/racecheck_$TAG.log | tail -8
if [ -n "$SPARSE" ]; then
echo "== memcheck (shared-footprint conv3_1 / fc1: work lists, difference epilogue, K-cell-sparse GEMM, row gathers, zero fill)"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 66 --print-limit 20 python -m pytest tests/test_gpu_fc1_shared.py tests/test_gpu_sparse.py -m gpu -q -x --timeout=1200 > $OUT/memcheck_sparse_$TAG.log 2>&1
echo "memcheck sparse exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|error" $OUT/memcheck_sparse_$TAG.log | tail -8
fi
