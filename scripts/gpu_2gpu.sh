#!/bin/bash
# 2-GPU visit: the default bench under torchrun exactly as the driver launches it (weak scaling, NCCL KPI gather).
TAG=${1:-r1g}
OUT=gpurun_out
mkdir -p $OUT
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus ${NG:-2} --steps 1 --warmup 3 ) > $OUT/${TAG}_bench_default_${NG:-2}gpu.json 2> $OUT/${TAG}_bench_default_${NG:-2}gpu.err
cut -c1-600 $OUT/${TAG}_bench_default_${NG:-2}gpu.json; tail -n 5 $OUT/${TAG}_bench_default_${NG:-2}gpu.err
