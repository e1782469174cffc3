#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -k "linear_tc or vr_scores_tc" -q -s --timeout 300 2>&1 | tail -45 > gpurun_out/pytest_tc.log
grep -E "linear_tc|passed|failed|Error" gpurun_out/pytest_tc.log | tail -40
if grep -q "failed\|rror" gpurun_out/pytest_tc.log; then echo "TC TESTS FAILED"; tail -30 gpurun_out/pytest_tc.log; fi
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/bench_full_f16x3.log | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print({k: d[k] for k in ('value','ms_per_step','phases_ms_per_step','corpus_encode','e2e')})
"
