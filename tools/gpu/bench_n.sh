#!/bin/bash
# usage: tools/gpu/bench_n.sh N [extra bench args]  -- one bench.py run on N GPUs of this box, JSON line to gpurun_out/
N=$1; shift
mkdir -p gpurun_out
tag=$(echo "n${N}$*" | tr -d ' -')
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 "$@" > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29871 bench.py --gpus $N "$@" > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
fi
tail -2 gpurun_out/bench_${tag}.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${tag}.json").read().strip().splitlines()[-1])
    print("N=%d value %.0f q/s  %.2f ms/step  e2e %s  exchange %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d.get("e2e", {}).get("value"), d["config"].get("exchange")))
    print({k: round(v, 3) for k, v in d["phases_ms_per_step"].items()})
except Exception as e:
    print("no JSON line:", e)
PY
