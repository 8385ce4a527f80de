#!/bin/bash
# GPU-box visit r1f: full GPU suite, smoke, default bench (FBMPC 8192/GPU), AB configs[1], reference arm,
# launch list of a small default-controller run.  bash scripts/gpu_r1f.sh [tag]
TAG=${1:-r1f}
OUT=gpurun_out
mkdir -p $OUT
( time timeout 1200 python -m pytest tests -m gpu -q -s ) > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
echo "smoke exit $?" >> $OUT/${TAG}_smoke.log
( time timeout 900 python bench.py ) > $OUT/${TAG}_bench_default_fb8192.json 2> $OUT/${TAG}_bench_default_fb8192.err
timeout 600 python bench.py --ctrl ab > $OUT/${TAG}_bench_ab4096.json 2> $OUT/${TAG}_bench_ab4096.err
timeout 400 python bench.py --impl reference --steps 1 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
SMALL="python bench.py --steps 1 --warmup 1 --cycle-steps 60 --batch 1184 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k5_lead -c 3 --csv --log-file $OUT/${TAG}_launches_k5s.csv $SMALL > $OUT/${TAG}_ncu_k5s.log 2>&1
tail -n 5 $OUT/${TAG}_pytest_gpu.log $OUT/${TAG}_smoke.log; cut -c1-400 $OUT/${TAG}_bench_*.json; tail -n 3 $OUT/${TAG}_bench_*.err
