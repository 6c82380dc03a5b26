#!/bin/bash
# round-2 GPU run 2 (2 GPUs): distributed parity tests + 2-GPU bench, two-chain schedule on/off
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_gpus.txt 2>&1
timeout 900 python -m pytest tests/test_dist.py -m gpu -x -q -k "two_gpus" > gpurun_out/r2_pytest_dist.log 2>&1; echo "pytest dist rc=$?"
tail -15 gpurun_out/r2_pytest_dist.log
run_bench() { # tag env...
  tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 \
    bench.py --gpus 2 --steps 3 --warmup 2 --e2e-steps 2 > gpurun_out/r2_bench_$tag.json 2> gpurun_out/r2_bench_$tag.err
  echo "bench $tag rc=$?"; tail -c 1500 gpurun_out/r2_bench_$tag.json | head -c 1500; echo
}
run_bench split DLAF_B200_SPLIT_CHAIN=1
run_bench nosplit DLAF_B200_SPLIT_CHAIN=0
grep -h "value\|Traceback\|Error" gpurun_out/r2_bench_*.err | tail -10
