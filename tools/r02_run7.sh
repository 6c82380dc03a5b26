#!/bin/bash
# round-2 GPU run 7 (8 GPUs): distributed parity on the 2x4 (BASELINE config C3) and 3x2 grids + bench on 2x4
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r7_gpus.txt 2>&1
timeout 600 python -m pytest tests/test_dist.py -m gpu -x -q -k "eight_gpus" > gpurun_out/r7_pytest_dist8.log 2>&1; echo "pytest dist8 rc=$?"
tail -3 gpurun_out/r7_pytest_dist8.log
run() { tag=$1; shift
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29731 \
    bench.py --gpus 8 --steps 3 --warmup 2 --e2e-steps 2 > gpurun_out/r7_bench_8gpu_$tag.json 2> gpurun_out/r7_bench_8gpu_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r7_bench_8gpu_$tag.json").read().strip().splitlines()[-1])
    c=d["roofline"]["critical_path_ms_last_step"]
    print("$tag", round(d["value"]), "GF/s", round(d["ms_per_step"],2), "ms | e2e", round(d["e2e"]["value"]) if d["e2e"] else None, "| bulk", round(d["roofline"]["kernel_ms_per_step"],1), "| potrf", round(c["diag_tile_potrf"],1), "wait_bulk", round(c["wait_bulk_and_diag_update"],1), "bcast", round(c["diag_bcast"],1), "trsm", round(c["wait_column_and_trsm"],1), "| residual", d["residual_max_diff_over_max_a"], "parity", d["oracle_parity"]["elementwise_vs_oracle_ok_all_ranks"] if d["oracle_parity"] else None)
except Exception as e:
    print("$tag failed", e)
PY
}
run split X=1
run split_r0 DLAF_B200_RESERVE_SMS=0
tail -2 gpurun_out/r7_bench_8gpu_split.err
