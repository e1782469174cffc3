#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_full_twopass.log | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print({k: d[k] for k in ('value','ms_per_step','phases_ms_per_step','e2e')}); print(d['clocks'], d['roofline']['frac'])
    else:
        print(l[-1500:])
"
