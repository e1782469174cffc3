#!/bin/bash
# launch lists (ncu gpu__time_duration per launch) of one timed search step and of one context-encoding batch
mkdir -p gpurun_out
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv"
timeout 600 $NCU --log-file gpurun_out/launches_step.csv python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-e2e --no-parity --no-gpu-reference --cuda-profiler > gpurun_out/launches_step.log 2>&1
timeout 300 $NCU --log-file gpurun_out/launches_encode.csv python tools/gpu/profile_encode.py > gpurun_out/launches_encode.log 2>&1
tail -2 gpurun_out/launches_step.log | cut -c1-200; tail -2 gpurun_out/launches_encode.log
python tools/gpu/profile_encode.py
# ... and of one training step (forward + backward, bsz 128)
timeout 300 $NCU --log-file gpurun_out/launches_train.csv python tools/gpu/profile_train.py > gpurun_out/launches_train.log 2>&1
tail -1 gpurun_out/launches_train.log | cut -c1-200
