#!/bin/bash
# source-level ncu of the 256-row pair-tile GEMMs at config 3: where do the epilogue warps wait?
O=gpurun_out/c15; mkdir -p $O
export KEEP_REP=1 NSB_BENCH_STREAMS=256 NSB_BENCH_R=6 NSB_BENCH_COMPUTE=bf16 NSB_BENCH_KV=bf16
bash tools/gpu_ncu_one.sh "gemm_tc_pair256" 6 c15/gemm_pair256_cfg3 > $O/gemm.log 2>&1
ncu -i $O/gemm_pair256_cfg3.ncu-rep --page source --csv > $O/gemm_pair256_cfg3_source.csv 2>/dev/null
cat $O/gemm_pair256_cfg3_summary.txt
rm -f $O/gemm_pair256_cfg3.ncu-rep
ls -la $O
