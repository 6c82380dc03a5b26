#!/bin/bash
# round-2 GPU run 12 (1 GPU): first hardware run of the inverse engine (tests + a measurement), triangular-solver tests after the refactoring
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_inverse_gpu.py -x -q > gpurun_out/r12_pytest_inverse.log 2>&1; echo "pytest inverse rc=$?"
tail -25 gpurun_out/r12_pytest_inverse.log
timeout 600 python -m pytest tests/test_triangular_gpu.py -x -q > gpurun_out/r12_pytest_trsm.log 2>&1; echo "pytest trsm rc=$?"
tail -3 gpurun_out/r12_pytest_trsm.log
timeout 300 python tools/bench_inverse.py --n 8192 --nb 512 > gpurun_out/r12_inverse_n8192.json 2> gpurun_out/r12_inverse_n8192.err; echo "bench rc=$?"
tail -c 1200 gpurun_out/r12_inverse_n8192.json; tail -3 gpurun_out/r12_inverse_n8192.err
timeout 300 python tools/bench_inverse.py --n 16384 --nb 512 --no-e2e > gpurun_out/r12_inverse_n16384.json 2> gpurun_out/r12_inverse_n16384.err; echo "bench rc=$?"
tail -c 1200 gpurun_out/r12_inverse_n16384.json; tail -3 gpurun_out/r12_inverse_n16384.err
