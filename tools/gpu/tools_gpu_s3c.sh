#!/bin/bash
# session-3 call C: full GPU suite, gradient accuracy report, default bench (N=1)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 400 python tests/perf_train_grad_accuracy.py > gpurun_out/train_grad_accuracy.txt 2>&1
tail -3 gpurun_out/train_grad_accuracy.txt | cut -c1-200
timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_s3c.log
python -c "
import json
d = json.loads(open('gpurun_out/bench_s3c.log').read()); print({k: d[k] for k in ('value','ms_per_step')}, {k: round(v, 2) for k, v in d['phases_ms_per_step'].items()}, d['e2e']['value'], d['clocks'], d['roofline']['frac'], d['roofline']['traffic'], d['gpu_launches'])
"
