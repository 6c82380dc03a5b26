#!/bin/bash
# round-2 GPU run 8 (1 GPU): cluster tile kernel v6, full GPU test suite, the default bench line, ncu launch list + full captures
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 200 tools/gpu_diag_tile_test > gpurun_out/r8_diag_tile.log 2>&1; echo "diag rc=$?"; tail -6 gpurun_out/r8_diag_tile.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r8_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r8_pytest.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/r8_clocks.csv &
SMI=$!
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r8_bench.json 2> gpurun_out/r8_bench.err; echo "bench rc=$?"
kill $SMI
timeout 600 python bench.py --n 16384 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r8_bench_n16384.json 2> gpurun_out/r8_bench_n16384.err; echo "bench16k rc=$?"
python - <<'PY'
import json
for f in ["r8_bench","r8_bench_n16384"]:
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), "GF/s", round(d["ms_per_step"],2), "ms e2e", round(d["e2e"]["value"]), d["e2e"].get("pageable_host"), "| roofline", round(d["roofline"]["achieved"],1), d["roofline"]["frac"], "| chain", {k: round(v,1) for k,v in d["roofline"]["critical_path_ms_last_step"].items()}, "| cusolver", d["gpu_library_reference"], "| cpu", d["cpu_baseline"] and d["cpu_baseline"]["value"])
    except Exception as e:
        print(f, "failed", e)
PY
# every launch of one factorization with its device time (shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r8_launches_n16384.csv \
  python bench.py --n 16384 --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-reference --no-check --parity-n 0 --e2e-steps 0 > gpurun_out/r8_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
# full captures of the hot kernels at the metric's own configuration (first launches of the first factorization: step k = 0)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"gemm_ozaki_i8_kernel|trsm_fused_f64_kernel|potrf_tile_cluster_kernel|split_i8_kernel" -c 8 \
  -o gpurun_out/r8_hot_kernels python bench.py --n 32768 --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-reference --no-check --parity-n 0 --e2e-steps 0 > gpurun_out/r8_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/r8_hot_kernels.ncu-rep
