#!/bin/bash
# round-2 GPU run 24 (1 GPU): end-to-end time vs upload chunk size (128 / 512 / 1024 MB; 256 = default measured before)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for mb in 512 128 1024; do
DLAF_B200_UPLOAD_CHUNK_MB=$mb timeout 100 python bench.py --steps 1 --warmup 3 --e2e-steps 3 --no-cpu-baseline --no-gpu-reference --next-n 0 --parity-n 0 --no-check > gpurun_out/r24_bench_chunk$mb.json 2> gpurun_out/r24_bench_chunk$mb.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r24_bench_chunk$mb.json").read().strip().splitlines()[-1])
    print("chunk $mb MB: value", round(d["value"]), "| e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],1), "ms")
except Exception as e:
    print("chunk $mb failed", e)
PY
done
