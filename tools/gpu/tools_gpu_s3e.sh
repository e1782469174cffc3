#!/bin/bash
# session-3 call E: model/sharded tests (external VR lists), ncu --set full of the secondary tensor-core kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_sharded.py tests/test_gpu_train.py -m gpu -q --timeout 600 --tb=short 2>&1 | tail -8
for k in span_probs_tc vr_rescore_tc linear_tc span_topk; do
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$k" -c 1 -o gpurun_out/prof_${k}_s3e -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --cuda-profiler > gpurun_out/ncu_$k.log 2>&1
tail -1 gpurun_out/ncu_$k.log | cut -c1-200
done
