#!/bin/bash
# round-2 GPU run 17 (2 GPUs): inverse / generalized->standard / triangular solver on a 2x1 grid at n=32768 and 16384
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for n in 32768 16384; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29771 \
  tools/bench_dist_next.py --grid 2x1 --matrix-size $n > gpurun_out/r17_next_2x1_n$n.json 2> gpurun_out/r17_next_2x1_n$n.err; echo "rc=$?"
tail -n 1 gpurun_out/r17_next_2x1_n$n.json | cut -c1-1500; tail -3 gpurun_out/r17_next_2x1_n$n.err
done
