#!/bin/bash
OUT=gpurun_out/${1:-sweep}; mkdir -p $OUT
for st in 8 12 16; do
  NSB_PAIR_STAGES=$st python tools/trace_step.py 3 > $OUT/trace_st$st.txt 2>&1
  echo "== pair stages $st: $(head -1 $OUT/trace_st$st.txt)"; grep "sum of pitches" $OUT/trace_st$st.txt; awk 'NR>=164 && NR<=178' $OUT/trace_st$st.txt | grep gemm_tc
done
