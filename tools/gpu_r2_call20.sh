#!/bin/bash
O=gpurun_out/c22; mkdir -p $O
export KEEP_REP=1 NSB_BENCH_STREAMS=256 NSB_BENCH_R=6 NSB_BENCH_COMPUTE=q8_0 NSB_BENCH_KV=f16
bash tools/gpu_ncu_one.sh "gemm_q8_pair256" 2 c22/q8pair > $O/q8pair.log 2>&1
ncu -i $O/q8pair.ncu-rep --page source --csv > $O/q8pair_source.csv 2>/dev/null
cat $O/q8pair_summary.txt
rm -f $O/q8pair.ncu-rep
timeout 600 python tools/trace_step.py 2 > $O/trace_cfg3_q8.txt 2>&1
grep -A12 "per kernel class" $O/trace_cfg3_q8.txt | cut -c1-170
