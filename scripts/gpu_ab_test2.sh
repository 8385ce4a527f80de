#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $OUT/t3_pytest_gpu.log 2>&1
B="python bench.py --steps 1 --warmup 2 --no-cpu-baseline"
run() { name=$1; shift; timeout 300 env "$@" $B --ctrl ab > $OUT/t3_${name}_ab.json 2> $OUT/t3_${name}_ab.err; timeout 300 env "$@" $B --ctrl fb > $OUT/t3_${name}_fb.json 2> $OUT/t3_${name}_fb.err; }
run std X=1
run fbg4 EEPACC_QP_FBGRID=4
run nt128 EEPACC_LIB=$PWD/eepacc_mpc_casadi_matlab_b200/libeepacc_nt128.so
for f in $OUT/t3_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d["value"]), round(d["e2e"]["value"]), d["kernel_ms_per_cycle"]["qp_ms"], d["latency_ms"]["p50"], d["latency_ms"]["p99"], d["latency_ms"]["max"], d["solver"], d["gpu_launches"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done > $OUT/t3_summary.txt 2>&1
SMALL="python bench.py --steps 1 --warmup 1 --cycle-steps 60 --batch 1184 --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k3r_qp -s 330 -c 2 -o $OUT/t3_k3r_full $SMALL > $OUT/t3_ncu_full.log 2>&1
