#!/bin/bash
# in-graph trace of a config 3 / config 5 step (bf16): per GEMM class prologue / main loop / epilogue of block 0
O=gpurun_out/c16; mkdir -p $O
NSB_BENCH_STREAMS=256 NSB_BENCH_R=6 NSB_BENCH_COMPUTE=bf16 NSB_BENCH_KV=bf16 timeout 600 python tools/trace_step.py 2 > $O/trace_cfg3_bf16.txt 2>&1
grep -B2 -A30 "per kernel class" $O/trace_cfg3_bf16.txt | cut -c1-200
grep "gemm_tc" $O/trace_cfg3_bf16.txt | sed -n 20,45p | cut -c1-160
