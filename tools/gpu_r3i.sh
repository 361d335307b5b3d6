#!/bin/bash
OUT=gpurun_out/${TAG:-r3i}; mkdir -p $OUT
timeout 500 python -m pytest tests -m gpu -q -x > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -12 $OUT/pytest.log | cut -c1-300
NSB_BENCH_STREAMS=256 NSB_BENCH_R=6 NSB_BENCH_COMPUTE=q8_0 NSB_BENCH_KV=f16 timeout 100 python tools/trace_step.py 2 > $OUT/trace_cfg3_q8.txt 2>&1; head -1 $OUT/trace_cfg3_q8.txt
NSB_BENCH_STREAMS=128 NSB_BENCH_R=0 timeout 100 python tools/trace_step.py 2 > $OUT/trace_cfg4_bf16.txt 2>&1; head -1 $OUT/trace_cfg4_bf16.txt
