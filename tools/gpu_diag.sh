#!/bin/bash
# diagnose a flaky / failing parity case under feature switches
OUT=gpurun_out/${1:-diag}; mkdir -p $OUT
K='test_streaming_parity_16bit_and_q8'
run() { echo "== $1"; for i in 1 2 3; do env $2 python -m pytest tests -m gpu -q -k "$K" 2>&1 | tail -4 | grep -E "passed|failed|FAILED" | tr '\n' ' '; echo; done; }
run "default" "X=1"
run "no split consumers" "NSB_NO_SPLIT_CONSUMERS=1"
run "no pair attention" "NSB_ATT_PAIR=0"
run "neither" "NSB_NO_SPLIT_CONSUMERS=1 NSB_ATT_PAIR=0"
run "no PDL" "NSB_NO_PDL=1"
python -m pytest tests -m gpu -q > $OUT/pytest_all.log 2>&1; tail -5 $OUT/pytest_all.log
