#!/bin/bash
# round-2 GPU run 22 (1 GPU): final tree: whole GPU test suite, default bench line, native-fp64-engine bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r22_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r22_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r22_bench.json 2> gpurun_out/r22_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r22_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "GF/s", round(d["ms_per_step"],1), "ms | e2e", round(d["e2e"]["value"]), "| residual", d["residual_max_diff_over_max_a"], "| parity", d["oracle_parity"]["elementwise_vs_oracle_ok_all_ranks"], "| roofline", round(d["roofline"]["achieved"]), round(d["roofline"]["frac"],3), "| cpu", d["cpu_baseline"]["value"])
n=d["next_rows"]; print("next_rows", {k:(round(v["value"]) if isinstance(v,dict) and "value" in v else v) for k,v in n.items()}, "hegst res", n["generalized_to_standard"]["max_LCLh_minus_A_over_max_A"])
PY
tail -2 gpurun_out/r22_bench.err
DLAF_B200_D_BULK=dmma timeout 600 python bench.py --steps 2 --warmup 3 --e2e-steps 2 --no-cpu-baseline --no-gpu-reference --next-n 0 --parity-n 0 > gpurun_out/r22_bench_native_fp64.json 2> gpurun_out/r22_bench_native_fp64.err; echo "bench native rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r22_bench_native_fp64.json").read().strip().splitlines()[-1])
print("native fp64 engine: value", round(d["value"]), "GF/s", round(d["ms_per_step"],1), "ms | e2e", round(d["e2e"]["value"]), "| residual", d["residual_max_diff_over_max_a"])
PY
