#!/bin/bash
# round-2 GPU run 1: int8 kernel v2 (both tile widths), full GPU test suite, default bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r1_gpu.txt 2>&1
timeout 300 tools/gpu_ozaki_test > gpurun_out/r1_ozaki_bn32.log 2>&1; echo "ozaki32 rc=$?"
DLAF_B200_OZAKI_BN=64 timeout 300 tools/gpu_ozaki_test > gpurun_out/r1_ozaki_bn64.log 2>&1; echo "ozaki64 rc=$?"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r1_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r1_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r1_bench.json
grep -E "ozaki_i8 |phase clocks|guard|FAILED" gpurun_out/r1_ozaki_bn32.log | head -30
