#!/bin/bash
OUT=gpurun_out/r3d; mkdir -p $OUT
timeout 400 python -m pytest tests -m gpu -q -x > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest.log
CFGS="0:0,256:2,128:3,128:4,256:97,208:97,160:97,112:97" timeout 200 python tools/gemm_large.py > $OUT/gemm_large_1792.txt 2>&1; echo "gemm_large rc=$?"; cat $OUT/gemm_large_1792.txt | tail -45
ROWS=896 CFGS="0:0,128:3,128:4,64:4,256:97,208:97,160:97,112:97" timeout 200 python tools/gemm_large.py > $OUT/gemm_large_896.txt 2>&1; echo "gemm_large896 rc=$?"; cat $OUT/gemm_large_896.txt | tail -45
timeout 200 python bench.py --steps 30 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cut -c1-400 $OUT/bench.json; grep -o '"roofline".*' $OUT/bench.json | cut -c1-900
NSB_BENCH_STREAMS=256 NSB_BENCH_R=6 NSB_BENCH_COMPUTE=f16 NSB_BENCH_KV=f16 timeout 200 python tools/trace_step.py 2 > $OUT/trace_cfg3_f16.txt 2>&1; head -1 $OUT/trace_cfg3_f16.txt; grep -A12 'per kernel class' $OUT/trace_cfg3_f16.txt | cut -c1-180
