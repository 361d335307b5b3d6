#!/bin/bash
# Q8_0 at 128 / 256 token rows (single steps, tools/ncu_step.py): fused dequantisation vs layer-ahead shadows
for cfg in "64 1" "128 1" "128 0" "16 6" "24 13"; do
  set -- $cfg
  for thr in 512 64; do
    echo -n "streams=$1 R=$2 thr=$thr: "
    NSB_Q8_PREDEQUANT_ROWS=$thr NSB_BENCH_STREAMS=$1 NSB_BENCH_R=$2 NSB_BENCH_COMPUTE=q8_0 NSB_BENCH_KV=f16 timeout 200 python tools/ncu_step.py 8 2>&1 | tail -1
  done
done
