#!/bin/bash
# usage: tools/gpu/ncu_kernel.sh <name> <kernel regex> <skip> <command...>  -- one ncu --set full capture into gpurun_out/
name=$1; regex=$2; skip=$3; shift 3
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$regex" -s $skip -c 1 -f -o gpurun_out/prof_$name "$@" > gpurun_out/prof_$name.log 2>&1
tail -2 gpurun_out/prof_$name.log | cut -c1-200
