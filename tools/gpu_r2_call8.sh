#!/bin/bash
O=gpurun_out/c8; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "overlap or two_steps or split_step or graph_replay or batch_push or api_edge or many_streams or serve" ) > $O/pytest.log 2>&1
tail -4 $O/pytest.log
NSB_DECODE_OVERLAP=1 timeout 300 python tools/trace_overlap.py 4 > $O/trace_overlap20.txt 2>&1
NSB_BENCH_STREAMS=128 NSB_BENCH_R=0 NSB_DECODE_OVERLAP=1 timeout 300 python tools/trace_overlap.py 4 > $O/trace_overlap20_cfg4.txt 2>&1
NSB_DECODE_OVERLAP=1 NSB_DECODE_CTAS=16 timeout 300 python tools/trace_overlap.py 4 > $O/trace_overlap16.txt 2>&1
( time timeout 900 python bench.py ) > $O/bench.json 2> $O/bench.err
tail -c 300 $O/bench.err
NSB_DECODE_OVERLAP=0 timeout 300 python bench.py --only-headline --no-cpu-baseline > $O/bench_no_overlap.json 2> $O/bench_no_overlap.err
for f in $O/trace_*.txt; do echo "== $f"; grep -E "^#|^decode|^step" $f | cut -c1-200; done
