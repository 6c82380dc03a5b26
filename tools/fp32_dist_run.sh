#!/bin/bash
# fp32 (tcgen05 3xTF32 path) on a 2-rank grid: parity tests + miniapp timing. usage: tools/fp32_dist_run.sh
timeout 900 python -m pytest tests/test_dist.py -x -q -m gpu 2>&1 | tail -3
run_mini() {  # $1 rows $2 cols $3 n $4 extra
  for r in 0 1; do
    RANK=$r WORLD_SIZE=2 LOCAL_RANK=$r MASTER_PORT=29811 DLAF_B200_RENDEZVOUS=/tmp/rdv_$1x$2_$3$5 $4 \
      timeout 300 ./miniapp/miniapp_cholesky --matrix-size $3 --block-size 512 --grid-rows $1 --grid-cols $2 \
      --type s --nruns 3 --nwarmups 1 > gpurun_out/mini_s_$1x$2_$3$5_r$r.log 2>&1 &
  done
  wait
  echo "== fp32 $1x$2 n=$3 $5"; grep -h "GFlop" gpurun_out/mini_s_$1x$2_$3$5_r0.log | tail -3
}
run_mini 2 1 32768 "env" tc
run_mini 1 2 32768 "env" tc
run_mini 2 1 32768 "env DLAF_B200_S_SIMT=1" simt
