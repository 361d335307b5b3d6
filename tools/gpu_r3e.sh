#!/bin/bash
OUT=gpurun_out/r3e; mkdir -p $OUT
timeout 400 python -m pytest tests -m gpu -q -x > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest.log
NSB_BENCH_STREAMS=256 NSB_BENCH_R=6 NSB_BENCH_COMPUTE=f16 NSB_BENCH_KV=f16 timeout 200 python tools/trace_step.py 2 > $OUT/trace_cfg3_f16.txt 2>&1; head -1 $OUT/trace_cfg3_f16.txt; grep -A9 'per kernel class' $OUT/trace_cfg3_f16.txt | cut -c1-180
NSB_BENCH_STREAMS=64 NSB_BENCH_R=13 timeout 200 python tools/trace_step.py 2 > $OUT/trace_cfg5_bf16.txt 2>&1; head -1 $OUT/trace_cfg5_bf16.txt; grep -A7 'per kernel class' $OUT/trace_cfg5_bf16.txt | cut -c1-180
NSB_BENCH_STREAMS=256 NSB_BENCH_R=6 NSB_BENCH_COMPUTE=q8_0 NSB_BENCH_KV=f16 timeout 200 python tools/trace_step.py 2 2>&1 | head -1
export NSB_BENCH_STREAMS=256 NSB_BENCH_R=6 NSB_BENCH_COMPUTE=f16 NSB_BENCH_KV=f16 KEEP_REP=1
timeout 400 bash tools/gpu_ncu_one.sh attention_mma 1 r3e_attn_cfg3
