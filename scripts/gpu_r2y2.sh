#!/bin/bash
# K6 launch-shape sweep (vehicles per warp x vehicles in flight per SM), then its GPU tests
OUT=gpurun_out; mkdir -p $OUT
: > $OUT/r2y2_k6_sweep.txt
for cfg in "32 64" "8 64" "4 64" "2 64" "8 32" "8 128" "4 32" "16 128"; do
  set -- $cfg
  echo "== lanes $1 inflight/SM $2" >> $OUT/r2y2_k6_sweep.txt
  EEPACC_K6_LANES=$1 EEPACC_K6_INFLIGHT=$2 timeout 60 python scripts/gpu_ocp_probe.py fb 8192 12 quick 2>&1 | grep "k QP/s" >> $OUT/r2y2_k6_sweep.txt
done
for cfg in "8 64" "4 64" "4 32"; do
  set -- $cfg
  echo "== AB lanes $1 inflight/SM $2" >> $OUT/r2y2_k6_sweep.txt
  EEPACC_K6_LANES=$1 EEPACC_K6_INFLIGHT=$2 timeout 60 python scripts/gpu_ocp_probe.py ab 8192 12 quick 2>&1 | grep "k QP/s" >> $OUT/r2y2_k6_sweep.txt
done
echo "== FB 65536 lanes 8 inflight 64" >> $OUT/r2y2_k6_sweep.txt
timeout 100 python scripts/gpu_ocp_probe.py fb 65536 6 quick 2>&1 | grep "k QP/s" >> $OUT/r2y2_k6_sweep.txt
cat $OUT/r2y2_k6_sweep.txt
timeout 200 python -m pytest tests/test_gpu_ocp.py -x -q -m gpu -s > $OUT/r2y2_pytest_ocp.log 2>&1; echo "pytest rc=$?" >> $OUT/r2y2_pytest_ocp.log
grep "structured vs dense" $OUT/r2y2_pytest_ocp.log | head -2; tail -3 $OUT/r2y2_pytest_ocp.log
