#!/bin/bash
# in-graph traces of the other BASELINE configs -> gpurun_out/<tag>/trace_<cfg>.txt (summary part only)
TAG=${1:-tr}; OUT=gpurun_out/$TAG; mkdir -p $OUT
run() { env $2 timeout 600 python tools/trace_step.py 2 > $OUT/trace_$1.txt 2>&1; echo "== $1 rc=$?"; grep -A40 "per kernel class" $OUT/trace_$1.txt | cut -c1-200; grep "decode:" $OUT/trace_$1.txt; head -1 $OUT/trace_$1.txt; }
run cfg3_f16  "NSB_BENCH_STREAMS=256 NSB_BENCH_R=6 NSB_BENCH_COMPUTE=f16 NSB_BENCH_KV=f16"
run cfg3_q8   "NSB_BENCH_STREAMS=256 NSB_BENCH_R=6 NSB_BENCH_COMPUTE=q8_0 NSB_BENCH_KV=f16"
run cfg5_bf16 "NSB_BENCH_STREAMS=64 NSB_BENCH_R=13 NSB_BENCH_COMPUTE=bf16 NSB_BENCH_KV=bf16"
run cfg4_bf16 "NSB_BENCH_STREAMS=128 NSB_BENCH_R=0 NSB_BENCH_COMPUTE=bf16 NSB_BENCH_KV=bf16"
