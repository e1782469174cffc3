#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:vr_scores_tc -c 1 -o gpurun_out/prof_vr_packed -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --cuda-profiler > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-300
