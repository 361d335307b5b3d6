#!/bin/bash
# Lean GPU call: parity suite, bench line, in-graph trace. Outputs under gpurun_out/<tag>/
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; RC=$?; echo "pytest rc=$RC"; tail -15 $OUT/pytest.log
python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cut -c1-700 $OUT/bench.json
python tools/trace_step.py 3 > $OUT/trace.txt 2>&1; echo "trace rc=$?"; tail -22 $OUT/trace.txt
if [ -n "$EXTRA" ]; then bash -c "$EXTRA"; fi
