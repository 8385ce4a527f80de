#!/bin/bash
# A/B of the smem-diet kernel: parity tests, then short bench lines under different capacities / variants
OUT=gpurun_out; mkdir -p $OUT
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $OUT/t2_pytest_gpu.log 2>&1
B="python bench.py --steps 1 --warmup 2 --no-cpu-baseline"
run() { name=$1; shift; timeout 300 env "$@" $B --ctrl ab > $OUT/t2_${name}_ab.json 2> $OUT/t2_${name}_ab.err; timeout 300 env "$@" $B --ctrl fb > $OUT/t2_${name}_fb.json 2> $OUT/t2_${name}_fb.err; }
run std X=1
run fc40 EEPACC_QP_FC=40
run fc39 EEPACC_QP_FC=39
run nofc2 EEPACC_QP_FC2=0
run nt128 EEPACC_LIB=$PWD/eepacc_mpc_casadi_matlab_b200/libeepacc_nt128.so
for f in $OUT/t2_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d["value"]), round(d["e2e"]["value"]), d["kernel_ms_per_cycle"]["qp_ms"], d["latency_ms"]["p50"], d["latency_ms"]["p99"], d["latency_ms"]["max"], d["solver"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done > $OUT/t2_summary.txt 2>&1
