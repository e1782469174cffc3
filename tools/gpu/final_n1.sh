#!/bin/bash
# final single-GPU call: full GPU suite, smoke(), default bench, launch list of one step
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -6 > gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_final.log
python -c "
import json
d = json.loads(open('gpurun_out/bench_final.log').read()); print({k: d[k] for k in ('value','ms_per_step')}, {k: round(v, 2) for k, v in d['phases_ms_per_step'].items()}, d['e2e']['value'], d['clocks'], round(d['roofline']['frac'], 3), d['roofline']['traffic'], d['gpu_launches'], d['cpu_baseline']['value'])
"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --cuda-profiler > gpurun_out/ncu_launches.log 2>&1
tail -1 gpurun_out/ncu_launches.log | cut -c1-200
