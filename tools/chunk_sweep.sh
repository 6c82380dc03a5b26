#!/bin/bash
# sweep DLAF_B200_BULK_CHUNKS for the device-resident bench
for c in 1 2 4 8; do
  for n in 16384 32768; do
    DLAF_B200_BULK_CHUNKS=$c timeout 300 python bench.py --n $n --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 --no-check 2>/dev/null > /tmp/sweep.json
    python - <<PY
import json
d=json.load(open('/tmp/sweep.json')); print("chunks $c n $n", round(d["value"]), round(d["ms_per_step"],2))
PY
  done
done
