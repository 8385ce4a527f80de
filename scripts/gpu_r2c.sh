#!/bin/bash
# A/B of k3r variants: e3 (cached ratio test + warp-synchronous wavefronts/Cholesky/classify), mb3 (= e3 at 3 CTAs/SM), phase timers
OUT=gpurun_out; mkdir -p $OUT
D=$PWD/eepacc_mpc_casadi_matlab_b200
for v in e3 mb3; do
  EEPACC_LIB=$D/libeepacc_$v.so timeout 300 python scripts/gpu_parity_quick.py fb 64 200 > $OUT/r2c_parity_$v.txt 2>&1
  EEPACC_LIB=$D/libeepacc_$v.so timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/r2c_bench_$v.json 2> $OUT/r2c_bench_$v.err
done
EEPACC_LIB=$D/libeepacc_mb3.so timeout 600 python bench.py --ctrl ab --steps 1 --warmup 1 --no-cpu-baseline > $OUT/r2c_bench_ab_mb3.json 2> $OUT/r2c_bench_ab_mb3.err
EEPACC_LIB=$D/libeepacc_e3.so timeout 600 python bench.py --ctrl ab --steps 1 --warmup 1 --no-cpu-baseline > $OUT/r2c_bench_ab_e3.json 2> $OUT/r2c_bench_ab_e3.err
export EEPACC_LIB=$D/libeepacc_e3t.so
for v in 0 3; do timeout 300 python scripts/gpu_phases_r.py fb 148 80 $v; done > $OUT/r2c_phases_fb.txt 2>&1
tail -2 $OUT/r2c_parity_*.txt; for f in $OUT/r2c_bench_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['latency_ms'], d['solver'])"; done; cat $OUT/r2c_phases_fb.txt
