#!/bin/bash
# round-2 GPU run 14 (1 GPU): first hardware run of the HEGST engine (tests + measurement); inverse tests after the refactoring; larger POTRI
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hegst_gpu.py -x -q > gpurun_out/r14_pytest_hegst.log 2>&1; echo "pytest hegst rc=$?"
tail -25 gpurun_out/r14_pytest_hegst.log
timeout 600 python -m pytest tests/test_inverse_gpu.py -x -q > gpurun_out/r14_pytest_inverse.log 2>&1; echo "pytest inverse rc=$?"
tail -3 gpurun_out/r14_pytest_inverse.log
for n in 8192 16384; do
timeout 300 python tools/bench_hegst.py --n $n --nb 512 > gpurun_out/r14_hegst_n$n.json 2> gpurun_out/r14_hegst_n$n.err; echo "bench rc=$?"
tail -c 1000 gpurun_out/r14_hegst_n$n.json; tail -3 gpurun_out/r14_hegst_n$n.err
done
timeout 300 python tools/bench_inverse.py --n 32768 --nb 512 --no-e2e --steps 2 > gpurun_out/r14_inverse_n32768.json 2> gpurun_out/r14_inverse_n32768.err; echo "bench rc=$?"
tail -c 1000 gpurun_out/r14_inverse_n32768.json; tail -3 gpurun_out/r14_inverse_n32768.err
