#!/bin/bash
# round-2 GPU run 19 (1 GPU): ring of digit-plane slots (host pipelining): POTRF tests, e2e with ring 16 vs ring 2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_potrf_gpu.py tests/test_ozaki_gpu.py tests/test_miniapp_gpu.py -x -q > gpurun_out/r19_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r19_pytest.log
for ring in 16 2; do
DLAF_B200_OZAKI_RING=$ring timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-reference --next-n 0 --parity-n 0 > gpurun_out/r19_bench_ring$ring.json 2> gpurun_out/r19_bench_ring$ring.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r19_bench_ring$ring.json").read().strip().splitlines()[-1])
print("ring $ring: value", round(d["value"]), "GF/s", round(d["ms_per_step"],1), "ms | e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],1), "ms res", d["e2e"].get("residual"), "| pageable", d["e2e"].get("pageable_host"))
PY
tail -2 gpurun_out/r19_bench_ring$ring.err
done
