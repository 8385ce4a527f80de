#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python scripts/gpu_parity_quick.py fb 64 200 > $OUT/r2g_parity_fb.txt 2>&1
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --parity-scenarios 0 > $OUT/r2g_bench_fb.json 2> $OUT/r2g_bench_fb.err
timeout 600 python bench.py --ctrl ab --steps 1 --warmup 1 --no-cpu-baseline --parity-scenarios 0 > $OUT/r2g_bench_ab.json 2> $OUT/r2g_bench_ab.err
tail -1 $OUT/r2g_parity_fb.txt | cut -c1-400
for f in $OUT/r2g_bench_fb.json $OUT/r2g_bench_ab.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['value'], d['e2e']['value'], d['latency_ms']['p50'], d['latency_ms']['p99'], d['latency_ms']['max'], d['kernel_ms_per_cycle'])"; done
