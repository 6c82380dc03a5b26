#!/bin/bash
# usage: tools/scale_run.sh "<list of gpu counts>"  — runs bench.py for each count on this box
for n in $1; do
  if [ "$n" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 2 --warmup 2 --no-cpu-baseline --e2e-steps 1 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2970$n bench.py --gpus $n --steps 2 --warmup 2 --no-cpu-baseline --e2e-steps 1 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  fi
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_$n.json").read().strip().splitlines()[-1])
    print("GPUS $n:", round(d["value"]), "GF/s", round(d["ms_per_step"],1), "ms  e2e", d["e2e"] and round(d["e2e"]["value"]), d["config"]["workload"])
except Exception as e:
    print("GPUS $n: FAILED", e); print(open("gpurun_out/scale_$n.err").read()[-1500:])
PY
done
