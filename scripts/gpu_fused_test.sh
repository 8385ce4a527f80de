#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k fused ) > $OUT/t5_pytest_fused.log 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $OUT/t5_pytest_gpu.log 2>&1
timeout 400 python bench.py --ctrl ab --steps 1 --warmup 2 --no-cpu-baseline > $OUT/t5_fused_ab.json 2> $OUT/t5_fused_ab.err
timeout 400 python bench.py --ctrl fb --steps 1 --warmup 2 --no-cpu-baseline > $OUT/t5_fused_fb.json 2> $OUT/t5_fused_fb.err
tail -3 $OUT/t5_pytest_fused.log; tail -3 $OUT/t5_pytest_gpu.log; cut -c1-400 $OUT/t5_fused_ab.json; tail -3 $OUT/t5_fused_ab.err
