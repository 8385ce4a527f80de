#!/bin/bash
# round-2 final evidence pass on the committed tree: GPU suite, smoke, default bench (+ AB, reference arm), ncu launch list and
# one --set full capture of the top kernel.  Numbers printed under ncu are never bench values.
TAG=${1:-r2z}
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -q -m gpu > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
timeout 900 python bench.py > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err; echo "rc=$?" >> $OUT/${TAG}_bench_default.err
timeout 600 python bench.py --impl reference > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
timeout 600 python bench.py --ctrl ab --steps 2 --warmup 1 > $OUT/${TAG}_bench_ab4096.json 2> $OUT/${TAG}_bench_ab4096.err
SMALL="python bench.py --steps 1 --warmup 1 --cycle-steps 60 --no-cpu-baseline --parity-scenarios 0"
timeout 300 $SMALL > $OUT/${TAG}_small_fb.json 2> $OUT/${TAG}_small_fb.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 400 --csv --log-file $OUT/${TAG}_launches_fb.csv $SMALL > $OUT/${TAG}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k3r_qp -s 400 -c 2 -o $OUT/${TAG}_k3r_fb_full $SMALL > $OUT/${TAG}_ncu_full.log 2>&1
tail -3 $OUT/${TAG}_pytest_gpu.log; tail -1 $OUT/${TAG}_smoke.log
for f in default ab4096 reference; do python -c "
import json
d=json.loads(open('$OUT/${TAG}_bench_$f.json').read().strip().splitlines()[-1]); print('$f', d['value'], d['e2e']['value'], d.get('latency_ms'), d.get('cpu_baseline'), d.get('parity'))"; done
ls -la $OUT | grep ${TAG}
