#!/bin/bash
# ncu --set full captures (one launch each) of the kernels of a search step and of a context-encoding batch -> gpurun_out/
mkdir -p gpurun_out
STEP="python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-e2e --no-parity --no-gpu-reference --cuda-profiler"
cap() {  # name regex skip command...
  name=$1; regex=$2; skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$regex" -s $skip -c 1 -f -o gpurun_out/r02_$name "$@" > gpurun_out/r02_$name.log 2>&1
  tail -1 gpurun_out/r02_$name.log | cut -c1-120
}
cap vr_filter_pair vr_filter_pair 0 $STEP
cap span_probs_tc span_probs_tc 0 $STEP
cap vr_rescore_tc vr_rescore_tc 0 $STEP
cap span_topk span_topk 0 $STEP
cap select_candidates select_candidates 0 $STEP
cap topk_rows topk_rows 0 $STEP
cap linear_tc_query linear_tc 3 $STEP
cap attention_ragged attention_ragged 0 $STEP
cap gather_rows16 gather_rows16 0 $STEP
cap attention_tc attention_tc 1 python tools/gpu/profile_encode.py
cap linear_tc_ctx linear_tc 0 python tools/gpu/profile_encode.py
cap add_layernorm add_layernorm 0 python tools/gpu/profile_encode.py
ls -la gpurun_out/r02_*.ncu-rep | awk '{print $5, $9}'
