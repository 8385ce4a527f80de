#!/bin/bash
# first GPU contact of the stage-structured solver K6: its tests, then short timing probes
OUT=gpurun_out; mkdir -p $OUT
timeout 500 python -m pytest tests/test_gpu_ocp.py -x -q -m gpu > $OUT/r2w_pytest_ocp.log 2>&1; echo "pytest rc=$?" >> $OUT/r2w_pytest_ocp.log
tail -5 $OUT/r2w_pytest_ocp.log
for c in fb ab; do
  timeout 300 python scripts/gpu_ocp_probe.py $c 8192 100 > $OUT/r2w_probe_$c.txt 2>&1; tail -3 $OUT/r2w_probe_$c.txt
done
timeout 200 python scripts/gpu_ocp_probe.py fb 65536 30 > $OUT/r2w_probe_fb_65536.txt 2>&1; tail -2 $OUT/r2w_probe_fb_65536.txt
