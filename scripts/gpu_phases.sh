#!/bin/bash
# phase timers of the reduced-space kernel (timers build) + per-step diagnostics
OUT=gpurun_out; mkdir -p $OUT
export EEPACC_LIB=$PWD/eepacc_mpc_casadi_matlab_b200/libeepacc_timers.so
for v in 0 1 2 3; do timeout 300 python scripts/gpu_phases_r.py ab 148 80 $v; done > $OUT/phases_ab.txt 2>&1
for v in 0 1 2 3; do timeout 300 python scripts/gpu_phases_r.py fb 148 80 $v; done > $OUT/phases_fb.txt 2>&1
unset EEPACC_LIB
timeout 300 python scripts/gpu_steps.py ab 4096 120 > $OUT/steps_ab.txt 2>&1
