#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -k "tc" -q --timeout 120 -x 2>&1 | tail -15 > gpurun_out/pytest_tc.log
tail -5 gpurun_out/pytest_tc.log
if grep -q "failed\|rror" gpurun_out/pytest_tc.log; then echo "TC TESTS FAILED"; cat gpurun_out/pytest_tc.log; exit 1; fi
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/bench_full_f16x3.log | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print({k: d[k] for k in ('value','ms_per_step','phases_ms_per_step','e2e')}); print(d['clocks'], d['roofline']['frac_of_peak_executed'])
"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:vr_scores_tc -c 1 -o gpurun_out/prof_vr_packed -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --cuda-profiler > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log | cut -c1-300
