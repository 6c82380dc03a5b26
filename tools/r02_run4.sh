#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 tools/gpu_diag_tile_test > gpurun_out/r4_diag_tile.log 2>&1; echo "diag tile test rc=$?"
cat gpurun_out/r4_diag_tile.log
timeout 600 python -m pytest tests/test_potrf_gpu.py tests/test_ozaki_gpu.py -m gpu -x -q > gpurun_out/r4_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r4_pytest.log
timeout 900 python bench.py --n 16384 --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-reference --parity-n 0 --e2e-steps 0 > gpurun_out/r4_bench_n16384.json 2> gpurun_out/r4_bench_n16384.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r4_bench_n16384.json"))
print(round(d["value"]), "GF/s", round(d["ms_per_step"],2), "ms", d["roofline"]["critical_path_ms_last_step"], d["residual_max_diff_over_max_a"])
PY
