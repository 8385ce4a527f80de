#!/bin/bash
# multi-GPU pass: N = $1 GPUs of one box
N=${1:-2}; OUT=gpurun_out; mkdir -p $OUT
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_gpu_edge_cases.py -q -m gpu -k two_devices > $OUT/r2m_two_device_test.log 2>&1; tail -2 $OUT/r2m_two_device_test.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 1 --warmup 1 > $OUT/r2m_bench_default_${N}gpu.json 2> $OUT/r2m_bench_default_${N}gpu.err
python -c "
import json
d=json.loads(open('$OUT/r2m_bench_default_${N}gpu.json').read().strip().splitlines()[-1]); print('N=$N', d['value'], d['e2e']['value'], d['latency_ms'], d['n_gpus'])"
B=$2
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/sweep_config4.py 871 $B tv > $OUT/r2m_sweep_tv_${N}gpu.jsonl 2> $OUT/r2m_sweep_tv_${N}gpu.err
cut -c1-400 $OUT/r2m_sweep_tv_${N}gpu.jsonl; tail -2 $OUT/r2m_sweep_tv_${N}gpu.err
