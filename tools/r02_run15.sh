#!/bin/bash
# round-2 GPU run 15 (1 GPU): workspace pool + panel GEMMs on the update engine: tests of the three sweeps, measurements at 3 sizes
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_hegst_gpu.py tests/test_inverse_gpu.py tests/test_triangular_gpu.py -x -q > gpurun_out/r15_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r15_pytest.log
for n in 8192 16384 32768; do
timeout 300 python tools/bench_hegst.py --n $n --nb 512 --steps 2 > gpurun_out/r15_hegst_n$n.json 2> gpurun_out/r15_hegst_n$n.err; echo "hegst rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/r15_hegst_n$n.json').read().strip().splitlines()[-1]);print('hegst',$n,round(d['value']),'GF/s',round(d['ms_device'],2),'ms res',d['residual_max_LCLh_minus_A_over_max_A'],'guard',d['guard_fallback_steps'],'ref',round(d['gpu_library_reference']['value_same_flop_model']))"
tail -2 gpurun_out/r15_hegst_n$n.err
timeout 300 python tools/bench_inverse.py --n $n --nb 512 --no-e2e --steps 2 > gpurun_out/r15_inverse_n$n.json 2> gpurun_out/r15_inverse_n$n.err; echo "inverse rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/r15_inverse_n$n.json').read().strip().splitlines()[-1]);print('potri',$n,round(d['value']),'GF/s',round(d['ms_device'],2),'ms res',d['residual_max_abs_invA_A_minus_I'],'guard',d['guard_fallback_steps'],'ref',round(d['gpu_library_reference']['value']) if d['gpu_library_reference'] and 'value' in d['gpu_library_reference'] else None)"
tail -2 gpurun_out/r15_inverse_n$n.err
done
