#!/bin/bash
OUT=gpurun_out/${TAG:-r3j}; mkdir -p $OUT
timeout 400 python -m pytest tests -m gpu -q -x -k "f32 or golden or invariance or dropin or large_batch" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest.log | cut -c1-300
timeout 200 python bench.py --steps 50 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$OUT/bench.json')); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'p50', d['p50_chunk_latency_ms'], d['breakdown']['subsampling'])"
NSB_BENCH_STREAMS=256 NSB_BENCH_R=6 NSB_BENCH_COMPUTE=f16 NSB_BENCH_KV=f16 timeout 100 python tools/trace_step.py 2 > $OUT/trace_cfg3_f16.txt 2>&1; head -1 $OUT/trace_cfg3_f16.txt; grep 'dwconv\|logmel\|stem' $OUT/trace_cfg3_f16.txt | tail -3 | cut -c1-150
