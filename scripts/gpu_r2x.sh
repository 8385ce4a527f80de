#!/bin/bash
# K6 after the latency work: GPU tests, timing probes, then a sanity pass of the default path (smoke + a short bench line)
OUT=gpurun_out; mkdir -p $OUT
timeout 330 python -m pytest tests/test_gpu_ocp.py -x -q -m gpu -s > $OUT/r2x_pytest_ocp.log 2>&1; echo "pytest rc=$?" >> $OUT/r2x_pytest_ocp.log
grep "structured vs dense" $OUT/r2x_pytest_ocp.log; tail -3 $OUT/r2x_pytest_ocp.log
for c in fb ab; do
  timeout 150 python scripts/gpu_ocp_probe.py $c 8192 40 quick > $OUT/r2x_probe_$c.txt 2>&1; tail -2 $OUT/r2x_probe_$c.txt
done
timeout 100 python __graft_entry__.py --smoke > $OUT/r2x_smoke.log 2>&1; tail -1 $OUT/r2x_smoke.log
timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --parity-scenarios 8 > $OUT/r2x_bench_default_short.json 2> $OUT/r2x_bench_default_short.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2x_bench_default_short.json"))
    print("default path:", d["value"], d["e2e"]["value"], d.get("parity"))
except Exception as e:
    print("bench line unreadable:", e)
PY
