#!/bin/bash
# Profile refresh: bench line, ncu launch list, in-graph trace, ncu --set full summaries of the hot kernels
TAG=${1:-prof}; OUT=gpurun_out/$TAG; mkdir -p $OUT
python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cut -c1-300 $OUT/bench.json
python tools/ncu_step.py 2 > $OUT/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/launches.csv python tools/ncu_step.py 2 > $OUT/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
python tools/trace_step.py 3 > $OUT/trace.txt 2>&1; echo "trace rc=$?"; grep "decode:" $OUT/trace.txt; tail -20 $OUT/trace.txt
bash tools/gpu_ncu_one.sh "gemm_tc" 9 $TAG/ncu_gemm
bash tools/gpu_ncu_one.sh "attention|conv_module|layernorm" 5 $TAG/ncu_misc
bash tools/gpu_ncu_one.sh "rnnt_decode|logmel|stem_conv0" 3 $TAG/ncu_decode
du -sh gpurun_out
