#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
export EEPACC_LIB=$PWD/eepacc_mpc_casadi_matlab_b200/libeepacc_timers.so
for v in 0 3; do timeout 300 python scripts/gpu_phases_r.py fb 148 80 $v; done > $OUT/r2j_phases_fb.txt 2>&1
for v in 1 2; do timeout 300 python scripts/gpu_phases_r.py ab 148 80 $v; done > $OUT/r2j_phases_ab.txt 2>&1
cat $OUT/r2j_phases_fb.txt $OUT/r2j_phases_ab.txt
