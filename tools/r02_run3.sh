#!/bin/bash
# round-2 GPU run 3: one-launch cluster kernel for the diagonal tile
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 tools/gpu_diag_tile_test > gpurun_out/r3_diag_tile.log 2>&1; echo "diag tile test rc=$?"
cat gpurun_out/r3_diag_tile.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r3_pytest.log
for n in 16384 32768; do
  timeout 900 python bench.py --n $n --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-reference --parity-n 0 --e2e-steps 0 > gpurun_out/r3_bench_n$n.json 2> gpurun_out/r3_bench_n$n.err; echo "bench $n rc=$?"
  DLAF_B200_POTRF_TILE=blocks timeout 900 python bench.py --n $n --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-reference --parity-n 0 --e2e-steps 0 --no-check > gpurun_out/r3_bench_n${n}_blocks.json 2> gpurun_out/r3_bench_n${n}_blocks.err; echo "bench $n blocks rc=$?"
done
python - <<'PY'
import json
for f in ["r3_bench_n16384","r3_bench_n16384_blocks","r3_bench_n32768","r3_bench_n32768_blocks"]:
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
        print(f, round(d["value"]), "GF/s", round(d["ms_per_step"],2), "ms", d["roofline"]["critical_path_ms_last_step"], d["residual_max_diff_over_max_a"])
    except Exception as e:
        print(f, "failed", e)
PY
