#!/bin/bash
# A/B: hand-overs through the global-work-area launch (default) vs the full-space kernel (EEPACC_QP_GWORK=0); cached opt-in; timers
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python scripts/gpu_parity_quick.py fb 64 200 > $OUT/r2f_parity_fb.txt 2>&1
timeout 300 python scripts/gpu_parity_quick.py ab 32 200 > $OUT/r2f_parity_ab.txt 2>&1
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --parity-scenarios 0 > $OUT/r2f_bench_fb.json 2> $OUT/r2f_bench_fb.err
EEPACC_QP_GWORK=0 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --parity-scenarios 0 > $OUT/r2f_bench_fb_nogwork.json 2> $OUT/r2f_bench_fb_nogwork.err
timeout 600 python bench.py --ctrl ab --steps 1 --warmup 1 --no-cpu-baseline --parity-scenarios 0 > $OUT/r2f_bench_ab.json 2> $OUT/r2f_bench_ab.err
export EEPACC_LIB=$PWD/eepacc_mpc_casadi_matlab_b200/libeepacc_timers.so
EEPACC_QP_MODE=1 timeout 300 python scripts/gpu_phases.py fb 148 4 > $OUT/r2f_phases_cold_fb.txt 2>&1
timeout 300 python scripts/gpu_phases_r.py fb 148 80 0 > $OUT/r2f_phases_fb.txt 2>&1
unset EEPACC_LIB
tail -1 $OUT/r2f_parity_fb.txt | cut -c1-400; tail -1 $OUT/r2f_parity_ab.txt | cut -c1-400
for f in $OUT/r2f_bench_fb.json $OUT/r2f_bench_fb_nogwork.json $OUT/r2f_bench_ab.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['value'], d['e2e']['value'], d['latency_ms']['p50'], d['latency_ms']['p99'], d['latency_ms']['max'], d['gpu_launches'])"; done
head -2 $OUT/r2f_phases_cold_fb.txt | cut -c1-900; cat $OUT/r2f_phases_fb.txt
