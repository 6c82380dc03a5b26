#!/bin/bash
# round-2 GPU run 6 (2 GPUs): why is the diagonal tile 2x slower in situ on grids? reservation / tiles-per-CTA / kernel variants
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
run() { tag=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29721 \
    bench.py --gpus 2 --steps 2 --warmup 1 --e2e-steps 0 --parity-n 0 --no-check > gpurun_out/r6_$tag.json 2> gpurun_out/r6_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r6_$tag.json").read().strip().splitlines()[-1])
    c=d["roofline"]["critical_path_ms_last_step"]
    print("$tag", round(d["value"]), "GF/s", round(d["ms_per_step"],2), "ms | bulk", round(d["roofline"]["kernel_ms_per_step"],1), "| potrf", round(c["diag_tile_potrf"],1), "wait_bulk", round(c["wait_bulk_and_diag_update"],1), "bcast", round(c["diag_bcast"],1), "trsm", round(c["wait_column_and_trsm"],1))
except Exception as e:
    print("$tag failed", e)
PY
}
run base X=1
run blocks DLAF_B200_POTRF_TILE=blocks
run reserve16 DLAF_B200_RESERVE_SMS=16
run reserve32 DLAF_B200_RESERVE_SMS=32
run tpc2 DLAF_B200_OZAKI_TPC=2
run tpc2_r16 DLAF_B200_OZAKI_TPC=2 DLAF_B200_RESERVE_SMS=16
