#!/bin/bash
# usage: gpu_var.sh name [ctrl...]  -- short bench of a variant library + (first call) QP operand dumps
OUT=gpurun_out; mkdir -p $OUT
V=$1; shift
[ "$V" != "std" ] && export EEPACC_LIB=$PWD/eepacc_mpc_casadi_matlab_b200/libeepacc_$V.so
for c in ${@:-ab}; do
  timeout 300 python bench.py --ctrl $c --steps 1 --warmup 2 --no-cpu-baseline > $OUT/var_${V}_$c.json 2> $OUT/var_${V}_$c.err
done
