#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 600 -x -k "index_grows or two_pass" 2>&1 | tail -5
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"span_topk" -c 1 -o gpurun_out/prof_span_topk -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --cuda-profiler > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-300
