#!/bin/bash
# round-2 GPU run 9 (4 GPUs): distributed tests incl. the triangular solver, BASELINE config C5 (ZPOTRF N=16384 nb=512, 2x2), cusolverMg comparator
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dist.py -m gpu -x -q -k "four_gpus" > gpurun_out/r9_pytest_dist4.log 2>&1; echo "pytest dist4 rc=$?"; tail -3 gpurun_out/r9_pytest_dist4.log
run() { tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29741 \
    bench.py --gpus 4 "$@" > gpurun_out/r9_bench_4gpu_$tag.json 2> gpurun_out/r9_bench_4gpu_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r9_bench_4gpu_$tag.json").read().strip().splitlines()[-1])
    print("$tag", d["metric"], round(d["value"]), "GF/s", round(d["ms_per_step"],2), "ms | e2e", round(d["e2e"]["value"]) if d["e2e"] else None, "| residual", d["residual_max_diff_over_max_a"], "parity", d["oracle_parity"]["elementwise_vs_oracle_ok_all_ranks"] if d["oracle_parity"] else None, "| roofline", d["roofline"]["kernel"][:40], d["roofline"]["achieved"], d["roofline"]["frac"])
except Exception as e:
    print("$tag failed", e)
PY
}
run c5_z --type z --n 16384 --nb 512 --steps 3 --warmup 3 --e2e-steps 2 --parity-n 4096
run c3_d --steps 3 --warmup 3 --e2e-steps 2
timeout 300 tools/cusolvermg_potrf_ref 32768 4 256 > gpurun_out/r9_cusolvermg_4gpu.log 2>&1; tail -2 gpurun_out/r9_cusolvermg_4gpu.log
timeout 300 tools/cusolvermg_potrf_ref 32768 2 256 > gpurun_out/r9_cusolvermg_2gpu.log 2>&1; tail -1 gpurun_out/r9_cusolvermg_2gpu.log
