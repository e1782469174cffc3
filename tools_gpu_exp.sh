#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --padded-corpus 2>&1 | tail -1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('PADDED', {k: d[k] for k in ('value','ms_per_step','phases_ms_per_step')}); print(d['clocks'], d['roofline']['frac_of_peak_executed'])
"
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:vr_scores_tc -c 1 -o gpurun_out/prof_vr_padded2 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --cuda-profiler --padded-corpus > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log | cut -c1-200
