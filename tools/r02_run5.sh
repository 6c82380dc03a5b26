#!/bin/bash
# round-2 GPU run 5 (4 GPUs): distributed parity tests (2x2, 1x4) + bench on the 2x2 grid
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
NG=${1:-4}
nvidia-smi -L > gpurun_out/r5_gpus.txt 2>&1
if [ "$NG" = "4" ]; then
  timeout 900 python -m pytest tests/test_dist.py -m gpu -x -q -k "two_gpus or four_gpus" > gpurun_out/r5_pytest_dist.log 2>&1; echo "pytest dist rc=$?"
  tail -4 gpurun_out/r5_pytest_dist.log
fi
run_bench() { # tag env...
  tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29711 \
    bench.py --gpus $NG --steps 3 --warmup 2 --e2e-steps 2 > gpurun_out/r5_bench_${NG}gpu_$tag.json 2> gpurun_out/r5_bench_${NG}gpu_$tag.err
  echo "bench $tag rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r5_bench_${NG}gpu_$tag.json"))
    print("$tag", round(d["value"]), "GF/s", round(d["ms_per_step"],2), "ms; e2e", round(d["e2e"]["value"]) if d["e2e"] else None, "; chain", d["roofline"]["critical_path_ms_last_step"], "; bulk ms", round(d["roofline"]["kernel_ms_per_step"],1), "; residual", d["residual_max_diff_over_max_a"], "; parity", d["oracle_parity"]["elementwise_vs_oracle_ok_all_ranks"] if d["oracle_parity"] else None)
except Exception as e:
    print("$tag failed", e)
PY
}
run_bench split DLAF_B200_SPLIT_CHAIN=1
run_bench nosplit DLAF_B200_SPLIT_CHAIN=0
tail -3 gpurun_out/r5_bench_${NG}gpu_split.err
