#!/bin/bash
TAG=${1:-r01d}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest gpu"; timeout 2000 python -m pytest tests -m gpu -q --timeout=900 > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -40 $OUT/pytest_$TAG.log
