#!/bin/bash
# capacity sweep of the reduced-space kernel (runtime knob): at 128 threads registers hold 2 CTAs/SM, so shared
# memory up to ~112 KB per CTA is free
OUT=gpurun_out; mkdir -p $OUT
for fc in 56 64 72; do
  for c in ab fb; do
    EEPACC_QP_FC=$fc timeout 300 python bench.py --ctrl $c --steps 1 --warmup 1 --no-cpu-baseline > $OUT/t7_fc${fc}_$c.json 2> $OUT/t7_fc${fc}_$c.err
  done
done
for f in $OUT/t7_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d["value"]), round(d["e2e"]["value"]), d["kernel_ms_per_cycle"]["qp_ms"], d["latency_ms"]["p50"], d["latency_ms"]["p99"], d["latency_ms"]["max"], d["solver"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done > $OUT/t7_summary.txt 2>&1
cat $OUT/t7_summary.txt
