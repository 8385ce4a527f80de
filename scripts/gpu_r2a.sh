#!/bin/bash
# round-2 first GPU pass: headline parity tests, the whole GPU suite, a short default bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
nproc >> gpurun_out/r2a_smi.txt
timeout 1500 python -m pytest tests/test_gpu_fb_headline.py -q -m gpu -s > gpurun_out/r2a_headline.log 2>&1
echo "headline rc=$?" >> gpurun_out/r2a_headline.log
timeout 1800 python -m pytest tests -q -m gpu --deselect tests/test_gpu_fb_headline.py > gpurun_out/r2a_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/r2a_suite.log
timeout 900 python bench.py --steps 2 --warmup 1 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?" >> gpurun_out/r2a_bench.err
tail -5 gpurun_out/r2a_headline.log; tail -15 gpurun_out/r2a_suite.log; cat gpurun_out/r2a_bench.json | head -c 1500
