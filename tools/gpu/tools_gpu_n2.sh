#!/bin/bash
# 2-GPU call: sharded-search tests + bench at N=2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -q --timeout 500 -x 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | grep '^{' | tail -1 > gpurun_out/bench_n2.log
python -c "
import json
d = json.loads(open('gpurun_out/bench_n2.log').read()); print('N=2', {k: d[k] for k in ('value','ms_per_step','phases_ms_per_step')}, d['e2e']['value'], d['clocks'])
"
