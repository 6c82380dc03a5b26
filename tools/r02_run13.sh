#!/bin/bash
# round-2 GPU run 13 (2 GPUs): distributed tests incl. the inverse engine on 2x1 and 1x2 grids
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_dist.py -m gpu -x -q > gpurun_out/r13_pytest_dist_2gpu.log 2>&1; echo "pytest dist rc=$?"
tail -30 gpurun_out/r13_pytest_dist_2gpu.log
