#!/bin/bash
# GPU-box visit for the use-case tests: bash scripts/gpu_usecases.sh [tag]
TAG=${1:-r1f}
OUT=gpurun_out
mkdir -p $OUT
( time timeout 1200 python -m pytest tests/test_gpu_use_cases.py -m gpu -q ) > $OUT/${TAG}_pytest_usecases.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest_usecases.log
tail -n 40 $OUT/${TAG}_pytest_usecases.log
