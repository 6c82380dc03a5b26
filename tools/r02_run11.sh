#!/bin/bash
# round-2 GPU run 11 (8 GPUs): BASELINE config C4 (fp32 POTRF N=65536 nb=1024, 2x4) and C3 (fp64 N=32768 nb=512, 2x4) with the final code
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
run() { tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29761 \
    bench.py --gpus 8 "$@" > gpurun_out/r11_bench_8gpu_$tag.json 2> gpurun_out/r11_bench_8gpu_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r11_bench_8gpu_$tag.json").read().strip().splitlines()[-1])
    print("$tag", d["metric"], round(d["value"]), "GF/s", round(d["ms_per_step"],2), "ms | e2e", round(d["e2e"]["value"]) if d["e2e"] else None, "| residual", d["residual_max_diff_over_max_a"], "gate", d["residual_gate_eps_n"], "parity", d["oracle_parity"]["elementwise_vs_oracle_ok_all_ranks"] if d["oracle_parity"] else None, "| chain", {k: round(v,1) for k,v in d["roofline"]["critical_path_ms_last_step"].items()}, "bulk", round(d["roofline"]["kernel_ms_per_step"],1))
except Exception as e:
    print("$tag failed", e)
PY
  tail -1 gpurun_out/r11_bench_8gpu_$tag.err
}
run c4_s --type s --matrix-size 65536 --block-size 1024 --steps 3 --warmup 3 --e2e-steps 2 --parity-n 8192
run c3_d --steps 3 --warmup 3 --e2e-steps 2
