#!/bin/bash
# every leg under its own timeout, outputs straight to files
OUT=gpurun_out/r3c; mkdir -p $OUT
timeout 300 python -m pytest tests -m gpu -q -x -k "gemm" > $OUT/pytest_gemm.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gemm.log
CFGS="0:0,256:2,128:3,256:97,208:97,160:97,112:97" timeout 200 python tools/gemm_large.py > $OUT/gemm_large_1792.txt 2>&1; echo "gemm_large rc=$?"; cat $OUT/gemm_large_1792.txt | tail -40
ROWS=896 CFGS="0:0,128:3,128:4,256:97,208:97,160:97,112:97" timeout 200 python tools/gemm_large.py > $OUT/gemm_large_896.txt 2>&1; echo "gemm_large896 rc=$?"; cat $OUT/gemm_large_896.txt | tail -40
NSB_SKIP=decode LANES=1,2 timeout 150 python tools/lanes_experiment.py 30 > $OUT/lanes_nodecode.txt 2>&1; echo "lanes rc=$?"; tail -3 $OUT/lanes_nodecode.txt
timeout 200 python bench.py --steps 30 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cut -c1-2600 $OUT/bench.json
