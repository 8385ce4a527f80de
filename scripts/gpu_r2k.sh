#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
D=$PWD/eepacc_mpc_casadi_matlab_b200
EEPACC_LIB=$D/libeepacc_k3nt128.so timeout 300 python scripts/gpu_parity_quick.py fb 32 40 > $OUT/r2k_parity_nt128.txt 2>&1
timeout 300 python bench.py --cycle-steps 1 --steps 3 --warmup 1 --no-cpu-baseline --parity-scenarios 0 > $OUT/r2k_cold_nt256.json 2> $OUT/r2k_cold_nt256.err
EEPACC_LIB=$D/libeepacc_k3nt128.so timeout 300 python bench.py --cycle-steps 1 --steps 3 --warmup 1 --no-cpu-baseline --parity-scenarios 0 > $OUT/r2k_cold_nt128.json 2> $OUT/r2k_cold_nt128.err
tail -1 $OUT/r2k_parity_nt128.txt | cut -c1-300
for f in nt256 nt128; do python -c "
import json
d=json.loads(open('$OUT/r2k_cold_$f.json').read().strip().splitlines()[-1]); print('$f cold step ms', d['ms_per_step'], d['value'])"; done
tail -2 $OUT/r2k_cold_nt128.err
