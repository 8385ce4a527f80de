#!/bin/bash
# phase timers (timers build): reduced-space kernel warm steps (FB, AB) and the full-space kernel's cold-start step
OUT=gpurun_out; mkdir -p $OUT
export EEPACC_LIB=$PWD/eepacc_mpc_casadi_matlab_b200/libeepacc_timers.so
for v in 0 3; do timeout 300 python scripts/gpu_phases_r.py fb 148 80 $v; done > $OUT/r2b_phases_fb.txt 2>&1
timeout 300 python scripts/gpu_phases_r.py ab 148 80 1 > $OUT/r2b_phases_ab.txt 2>&1
EEPACC_QP_MODE=1 timeout 300 python scripts/gpu_phases.py fb 148 8 > $OUT/r2b_phases_cold_fb.txt 2>&1
cat $OUT/r2b_phases_fb.txt $OUT/r2b_phases_cold_fb.txt
