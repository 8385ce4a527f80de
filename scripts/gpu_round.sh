#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench lines, ncu launch list + full capture of the dominant kernel.
# Usage (from the repo root on the box): bash scripts/gpu_round.sh [tag]
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > $OUT/${TAG}_gpu.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
timeout 600 python bench.py > $OUT/${TAG}_bench_ab.json 2> $OUT/${TAG}_bench_ab.err
timeout 600 python bench.py --ctrl fb --steps 1 > $OUT/${TAG}_bench_fb.json 2> $OUT/${TAG}_bench_fb.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
timeout 400 python bench.py --fused --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_bench_ab_fused.json 2> $OUT/${TAG}_bench_ab_fused.err
SMALL="python bench.py --steps 1 --warmup 1 --cycle-steps 60 --batch 1184 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file $OUT/${TAG}_launches.csv $SMALL > $OUT/${TAG}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k3r_qp -s 170 -c 2 -o $OUT/${TAG}_k3r_full $SMALL > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT | tail -15
