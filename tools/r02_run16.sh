#!/bin/bash
# round-2 GPU run 16 (4 GPUs): distributed tests (POTRF, solver, inverse, generalized -> standard) on 2x2 and 1x4 grids
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dist.py -m gpu -k four -x -q > gpurun_out/r16_pytest_dist_4gpu.log 2>&1; echo "pytest dist rc=$?"
tail -30 gpurun_out/r16_pytest_dist_4gpu.log
