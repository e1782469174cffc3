#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -k "packed" -q -s --timeout 300 2>&1 | tail -45 > gpurun_out/pytest_tc.log
grep -E "passed|failed|Error" gpurun_out/pytest_tc.log | tail -5
if grep -q "failed" gpurun_out/pytest_tc.log; then echo "TC TESTS FAILED"; tail -30 gpurun_out/pytest_tc.log; exit 1; fi
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/bench_full_f16x3.log | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print({k: d[k] for k in ('value','ms_per_step','phases_ms_per_step','e2e')}); print(d['roofline'])
"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:vr_scores_tc -c 1 -o gpurun_out/prof_vr_packed -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --cuda-profiler > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log | cut -c1-300
