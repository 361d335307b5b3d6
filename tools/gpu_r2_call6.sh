#!/bin/bash
O=gpurun_out/c6; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "not config2 and not encoder_error" ) > $O/pytest.log 2>&1
tail -4 $O/pytest.log
( time timeout 900 python bench.py ) > $O/bench.json 2> $O/bench.err
tail -c 300 $O/bench.err
NSB_BENCH_STREAMS=64 NSB_BENCH_R=13 timeout 300 python tools/trace_step.py 2 > $O/trace_cfg5.txt 2>&1
tail -22 $O/trace_cfg5.txt
