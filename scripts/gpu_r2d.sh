#!/bin/bash
# round-2 evidence pass: GPU suite, default bench (+ reference arm), ncu launch list and one --set full capture of the top kernel
TAG=${1:-r2d}
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -q -m gpu > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest_gpu.log
timeout 900 python bench.py --steps 2 --warmup 1 > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err; echo "rc=$?" >> $OUT/${TAG}_bench_default.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
SMALL="python bench.py --steps 1 --warmup 1 --cycle-steps 60 --no-cpu-baseline --parity-scenarios 0"
timeout 300 $SMALL > $OUT/${TAG}_small_fb.json 2> $OUT/${TAG}_small_fb.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 400 --csv --log-file $OUT/${TAG}_launches_fb.csv $SMALL > $OUT/${TAG}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k3r_qp -s 400 -c 2 -o $OUT/${TAG}_k3r_fb_full $SMALL > $OUT/${TAG}_ncu_full.log 2>&1
tail -3 $OUT/${TAG}_pytest_gpu.log; head -c 600 $OUT/${TAG}_bench_default.json; echo; cat $OUT/${TAG}_bench_reference.json | head -c 400; echo; cat $OUT/${TAG}_smoke.log | tail -2; ls -la $OUT | tail -12
