#!/bin/bash
OUT=gpurun_out/${TAG:-r3g}; mkdir -p $OUT
timeout 400 python -m pytest tests -m gpu -q -x > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest.log
timeout 200 python tools/trace_step.py 3 > $OUT/trace_cfg2.txt 2>&1; head -1 $OUT/trace_cfg2.txt; sed -n 2,36p $OUT/trace_cfg2.txt | cut -c1-120; grep -A14 'per kernel class' $OUT/trace_cfg2.txt | cut -c1-180
NSB_BENCH_STREAMS=256 NSB_BENCH_R=6 NSB_BENCH_COMPUTE=f16 NSB_BENCH_KV=f16 timeout 200 python tools/trace_step.py 2 > $OUT/trace_cfg3_f16.txt 2>&1; head -1 $OUT/trace_cfg3_f16.txt; grep 'stem\|logmel' $OUT/trace_cfg3_f16.txt | tail -2 | cut -c1-180
timeout 200 python bench.py --steps 30 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cut -c1-330 $OUT/bench.json
