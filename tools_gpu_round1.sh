#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_full_f16x3.log | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print({k: d[k] for k in ('value','ms_per_step','phases_ms_per_step','e2e', 'cpu_baseline')}); print(d['clocks'], d['roofline']['frac_of_peak_executed'])
"
