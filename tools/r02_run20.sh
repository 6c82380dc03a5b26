#!/bin/bash
# round-2 GPU run 20 (1 GPU): e2e with downloads deferred until the upload is complete vs both directions competing
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for defer in 1 0; do
DLAF_B200_HOST_DEFER_D2H=$defer timeout 600 python bench.py --steps 1 --warmup 3 --e2e-steps 4 --no-cpu-baseline --no-gpu-reference --next-n 0 --parity-n 0 > gpurun_out/r20_bench_defer$defer.json 2> gpurun_out/r20_bench_defer$defer.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r20_bench_defer$defer.json").read().strip().splitlines()[-1])
print("defer $defer: value", round(d["value"]), "GF/s", round(d["ms_per_step"],1), "ms | e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],1), "ms res", d["e2e"].get("residual"), "| pageable", round(d["e2e"]["pageable_host"]["ms_per_step"],1))
PY
tail -2 gpurun_out/r20_bench_defer$defer.err
done
timeout 300 python -m pytest tests/test_potrf_gpu.py -x -q -k "golden or sentinel or scalapack or host" > gpurun_out/r20_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r20_pytest.log
