#!/bin/bash
# round-2 GPU run 21 (2 GPUs): the driver's N=2 launch of the final tree (bench line with grid residual + oracle parity)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29781 \
  bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r21_bench_2gpu.json 2> gpurun_out/r21_bench_2gpu.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r21_bench_2gpu.json").read().strip().splitlines()[-1])
print("2 GPUs:", round(d["value"]), "GF/s", round(d["ms_per_step"],1), "ms | e2e", round(d["e2e"]["value"]), "| residual", d["residual_max_diff_over_max_a"], "| parity", d["oracle_parity"]["elementwise_vs_oracle_ok_all_ranks"], "| launches", d["gpu_launches"])
PY
tail -2 gpurun_out/r21_bench_2gpu.err
