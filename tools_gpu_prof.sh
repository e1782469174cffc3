#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -q --timeout 600 -x -k "two_pass or select or hi_err" 2>&1 | tail -5
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"vr_scores_tc_packed|vr_rescore" -c 2 -o gpurun_out/prof_vr_twopass -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --cuda-profiler > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r01c.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --cuda-profiler > gpurun_out/ncu_launches.log 2>&1
tail -1 gpurun_out/ncu_launches.log | cut -c1-200
