#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python scripts/gpu_parity_quick.py fb 64 200 > $OUT/r2h_parity_fb.txt 2>&1
timeout 300 python scripts/gpu_parity_quick.py ab 32 200 > $OUT/r2h_parity_ab.txt 2>&1
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --parity-scenarios 0 > $OUT/r2h_bench_fb.json 2> $OUT/r2h_bench_fb.err
timeout 600 python bench.py --ctrl ab --steps 1 --warmup 1 --no-cpu-baseline --parity-scenarios 0 > $OUT/r2h_bench_ab.json 2> $OUT/r2h_bench_ab.err
timeout 600 python bench.py --fused --steps 1 --warmup 1 --no-cpu-baseline --parity-scenarios 0 > $OUT/r2h_bench_fb_fused.json 2> $OUT/r2h_bench_fb_fused.err
timeout 600 python bench.py --ctrl ab --fused --steps 1 --warmup 1 --no-cpu-baseline --parity-scenarios 0 > $OUT/r2h_bench_ab_fused.json 2> $OUT/r2h_bench_ab_fused.err
tail -1 $OUT/r2h_parity_fb.txt | cut -c1-300; tail -1 $OUT/r2h_parity_ab.txt | cut -c1-300
for f in fb ab fb_fused ab_fused; do python -c "
import json
d=json.loads(open('$OUT/r2h_bench_$f.json').read().strip().splitlines()[-1]); print('$f', d['value'], d['e2e']['value'], d['latency_ms']['p50'], d['latency_ms']['p99'], d['latency_ms']['max'], d['kernel_ms_per_cycle']['qp_ms'], d.get('fused'))"; done
