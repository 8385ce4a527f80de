#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -q -m gpu > $OUT/r2e_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/r2e_pytest_gpu.log
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --parity-scenarios 0 > $OUT/r2e_bench_fb.json 2> $OUT/r2e_bench_fb.err
timeout 600 python bench.py --ctrl ab --steps 1 --warmup 1 --no-cpu-baseline --parity-scenarios 0 > $OUT/r2e_bench_ab.json 2> $OUT/r2e_bench_ab.err
timeout 600 python scripts/sweep_config4.py 60 1024,16384 tv,bl > $OUT/r2e_sweep_smoke.jsonl 2> $OUT/r2e_sweep_smoke.err
tail -4 $OUT/r2e_pytest_gpu.log; for f in $OUT/r2e_bench_fb.json $OUT/r2e_bench_ab.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['latency_ms']['p50'], d['latency_ms']['p99'], d['latency_ms']['max'], d['solver'])"; done; cat $OUT/r2e_sweep_smoke.jsonl | cut -c1-400; tail -3 $OUT/r2e_sweep_smoke.err
