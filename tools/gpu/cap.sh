#!/bin/bash
# usage: tools/gpu/cap.sh <tag> <kernel regex> <skip>   -- one ncu --set full capture of a kernel of a search step
name=$1; regex=$2; skip=$3
mkdir -p gpurun_out
STEP="python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-e2e --no-parity --no-gpu-reference --cuda-profiler"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$regex" -s $skip -c 1 -f -o gpurun_out/$name $STEP > gpurun_out/$name.log 2>&1
tail -1 gpurun_out/$name.log | cut -c1-120
