#!/bin/bash
# One GPU call: parity suite, bench, ncu launch list, in-graph trace, ncu --set full of the GEMM / attention kernels, other configs
TAG=${1:-r1s}
OUT=gpurun_out/$TAG
mkdir -p $OUT
NCU=1 bash tools/gpu_round.sh $TAG
python tools/trace_step.py 3 > $OUT/trace.txt 2>&1; echo "trace rc=$?"; tail -22 $OUT/trace.txt
bash tools/gpu_ncu_one.sh "gemm_tc" 8 $TAG/ncu_gemm
bash tools/gpu_ncu_one.sh "attention_kernel|conv_module|layernorm" 5 $TAG/ncu_misc
bash tools/gpu_configs.sh > $OUT/configs.txt 2>&1; cat $OUT/configs.txt
du -sh gpurun_out
