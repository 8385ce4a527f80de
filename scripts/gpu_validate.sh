#!/bin/bash
# Short GPU-box visit: parity tests, smoke, the three bench lines (no profiler).  bash scripts/gpu_validate.sh [tag]
TAG=${1:-r1e}
OUT=gpurun_out
mkdir -p $OUT
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
echo "smoke exit $?" >> $OUT/${TAG}_smoke.log
timeout 600 python bench.py > $OUT/${TAG}_bench_ab.json 2> $OUT/${TAG}_bench_ab.err
timeout 600 python bench.py --ctrl fb --steps 1 > $OUT/${TAG}_bench_fb.json 2> $OUT/${TAG}_bench_fb.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
tail -3 $OUT/${TAG}_pytest_gpu.log $OUT/${TAG}_smoke.log; cut -c1-300 $OUT/${TAG}_bench_*.json
