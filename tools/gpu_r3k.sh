#!/bin/bash
OUT=gpurun_out/${TAG:-r3k}; mkdir -p $OUT
timeout 300 python -m pytest tests -m gpu -q -x -k "large_batch" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest.log | cut -c1-300
export NSB_BENCH_STREAMS=256 NSB_BENCH_R=6 NSB_BENCH_COMPUTE=f16 NSB_BENCH_KV=f16
timeout 100 python tools/trace_step.py 3 2>&1 | grep "^# step"
NSB_FFDOWN_SPLIT=1 timeout 100 python tools/trace_step.py 3 2>&1 | grep "^# step"
NSB_BENCH_COMPUTE=q8_0 timeout 100 python tools/trace_step.py 3 2>&1 | grep "^# step"
