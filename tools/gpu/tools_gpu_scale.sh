#!/bin/bash
# multi-GPU bench at the GPU count of this box (argument), one JSON line per run into gpurun_out/bench_n<N>.log
mkdir -p gpurun_out
n=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 5 --warmup 3 2>&1 | grep '^{' | tail -1 > gpurun_out/bench_n$n.log
python -c "
import json
d = json.loads(open('gpurun_out/bench_n$n.log').read()); print('N=$n', {k: d[k] for k in ('value','ms_per_step')}, {k: round(v, 2) for k, v in d['phases_ms_per_step'].items()}, d['e2e']['value'], d['clocks'])
"
