#!/bin/bash
# GPU-box visit r1g (FBMPC evidence for the default bench): launch list + one --set full capture of k3r_qp_kernel
# on the FB controller (steady-state step of a 60-step cycle).  Numbers printed under ncu are never bench values.
TAG=${1:-r1g}
OUT=gpurun_out
mkdir -p $OUT
SMALL="python bench.py --ctrl fb --steps 1 --warmup 1 --cycle-steps 60 --batch 1184 --no-cpu-baseline"
timeout 300 $SMALL > $OUT/${TAG}_small_fb.json 2> $OUT/${TAG}_small_fb.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file $OUT/${TAG}_launches_fb.csv $SMALL > $OUT/${TAG}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k3r_qp -s 170 -c 2 -o $OUT/${TAG}_k3r_fb_full $SMALL > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT | tail -8
