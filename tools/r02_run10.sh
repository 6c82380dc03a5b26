#!/bin/bash
# round-2 GPU run 10 (4 GPUs): BASELINE config C5 (ZPOTRF N=16384 nb=512, 2x2) and the fp32 nb=1024 path of C4 at N=32768
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
run() { tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29751 \
    bench.py --gpus 4 "$@" > gpurun_out/r10_bench_4gpu_$tag.json 2> gpurun_out/r10_bench_4gpu_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r10_bench_4gpu_$tag.json").read().strip().splitlines()[-1])
    print("$tag", d["metric"], round(d["value"]), "GF/s", round(d["ms_per_step"],2), "ms | e2e", round(d["e2e"]["value"]) if d["e2e"] else None, "| residual", d["residual_max_diff_over_max_a"], "gate", d["residual_gate_eps_n"], "parity", d["oracle_parity"]["elementwise_vs_oracle_ok_all_ranks"] if d["oracle_parity"] else None, "| roofline", d["roofline"]["kernel"][:40], d["roofline"]["achieved"], d["roofline"]["frac"])
except Exception as e:
    print("$tag failed", e)
PY
  tail -2 gpurun_out/r10_bench_4gpu_$tag.err
}
run c5_z --type z --matrix-size 16384 --block-size 512 --steps 3 --warmup 3 --e2e-steps 2 --parity-n 4096
run s_nb1024 --type s --matrix-size 32768 --block-size 1024 --steps 3 --warmup 3 --e2e-steps 2 --parity-n 4096
