#!/bin/bash
# round-2 GPU run 18 (1 GPU): validation of the final tree: smoke(), the whole GPU test suite, the default bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r18_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r18_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r18_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r18_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r18_bench.json 2> gpurun_out/r18_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r18_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "GF/s", round(d["ms_per_step"],1), "ms | e2e", round(d["e2e"]["value"]), "| residual", d["residual_max_diff_over_max_a"], "| parity", d["oracle_parity"]["elementwise_vs_oracle_ok_all_ranks"], "| roofline", d["roofline"]["achieved"], d["roofline"]["frac"], "| cpu", d["cpu_baseline"]["value"], "| lib", d["gpu_library_reference"])
print("next_rows", json.dumps(d["next_rows"])[:1500])
PY
tail -3 gpurun_out/r18_bench.err
