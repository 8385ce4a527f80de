#!/bin/bash
# fused kernel: residency sweep (does a co-resident CTA slow its neighbours?) + phase split; inline vs noinline build
OUT=gpurun_out; mkdir -p $OUT
run() { name=$1; ctrl=$2; shift; shift; timeout 400 env "$@" python bench.py --ctrl $ctrl --steps 1 --warmup 1 --no-cpu-baseline > $OUT/t6_${name}_${ctrl}.json 2> $OUT/t6_${name}_${ctrl}.err; }
NOINL=EEPACC_LIB=$PWD/eepacc_mpc_casadi_matlab_b200/libeepacc_noinl.so
run inl_g148 ab EEPACC_CYCLE_GRID=148
run inl_g444 ab X=1
run noinl_g148 ab $NOINL EEPACC_CYCLE_GRID=148
run noinl_g296 ab $NOINL EEPACC_CYCLE_GRID=296
run noinl_g444 ab $NOINL
run noinl_g444 fb $NOINL
run noinl_perstep ab $NOINL EEPACC_FUSED=0
for f in $OUT/t6_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d["value"]), d.get("fused"))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done > $OUT/t6_summary.txt 2>&1
EEPACC_LIB=$PWD/eepacc_mpc_casadi_matlab_b200/libeepacc_noinl.so timeout 600 python -m pytest tests -m gpu -x -q > $OUT/t6_pytest_noinl.log 2>&1
