#!/bin/bash
OUT=gpurun_out/${TAG:-r3f}; mkdir -p $OUT
timeout 400 python -m pytest tests -m gpu -q -x > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest.log
NSB_BENCH_STREAMS=256 NSB_BENCH_R=6 NSB_BENCH_COMPUTE=f16 NSB_BENCH_KV=f16 timeout 200 python tools/trace_step.py 2 > $OUT/trace_cfg3_f16.txt 2>&1; head -1 $OUT/trace_cfg3_f16.txt; grep -A7 'per kernel class' $OUT/trace_cfg3_f16.txt | cut -c1-180
NSB_BENCH_STREAMS=64 NSB_BENCH_R=13 timeout 200 python tools/trace_step.py 2 > $OUT/trace_cfg5_bf16.txt 2>&1; head -1 $OUT/trace_cfg5_bf16.txt; grep -A5 'per kernel class' $OUT/trace_cfg5_bf16.txt | cut -c1-180
timeout 200 python bench.py --steps 30 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cut -c1-330 $OUT/bench.json; grep -o '"breakdown".*' $OUT/bench.json | cut -c1-500
