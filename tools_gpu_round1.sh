#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --cuda-profiler > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:vr_scores_tc -c 2 -o gpurun_out/prof_vr_r01 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --cuda-profiler > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-300
ls -la gpurun_out
