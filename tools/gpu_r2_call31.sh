#!/bin/bash
for cfg in "1 1" "8 1" "32 1" "48 1"; do
  set -- $cfg
  for thr in 512 1; do
    echo -n "streams=$1 R=$2 thr=$thr: "
    NSB_Q8_PREDEQUANT_ROWS=$thr NSB_BENCH_STREAMS=$1 NSB_BENCH_R=$2 NSB_BENCH_COMPUTE=q8_0 NSB_BENCH_KV=f16 timeout 200 python tools/ncu_step.py 6 2>&1 | tail -1 | cut -c1-120
  done
done
