#!/bin/bash
# GPU-box visit r1i: the time-parallel K5s kernel -- its tests, the probe plain and under an ncu launch list
# (duration + DRAM bytes per launch), then the default bench line.
TAG=${1:-r1i}
OUT=gpurun_out
mkdir -p $OUT
( timeout 600 python -m pytest tests/test_gpu_use_cases.py -m gpu -q -k "synthesis" ) > $OUT/${TAG}_pytest_k5s.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest_k5s.log
timeout 300 python scripts/k5s_probe.py > $OUT/${TAG}_k5s_probe.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_write.sum,dram__bytes_read.sum --clock-control none -k regex:k5_lead --csv --log-file $OUT/${TAG}_k5s_launches.csv python scripts/k5s_probe.py > $OUT/${TAG}_k5s_ncu.log 2>&1
( time timeout 900 python bench.py ) > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err
tail -n 3 $OUT/${TAG}_pytest_k5s.log; cat $OUT/${TAG}_k5s_probe.txt; cut -c1-300 $OUT/${TAG}_bench_default.json
