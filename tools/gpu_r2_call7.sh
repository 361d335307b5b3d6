#!/bin/bash
O=gpurun_out/c7; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "not config2 and not encoder_error" ) > $O/pytest.log 2>&1
tail -4 $O/pytest.log
( time timeout 900 python bench.py ) > $O/bench.json 2> $O/bench.err
tail -c 300 $O/bench.err
NSB_LN_ROWS_MIN=100000 timeout 300 python bench.py --config 3 --only-headline --no-cpu-baseline > $O/bench_cfg3_oldln.json 2> $O/bench_cfg3_oldln.err
NSB_CONV_TB=14 timeout 300 python bench.py --config 5 --only-headline --no-cpu-baseline > $O/bench_cfg5_tb14.json 2> $O/bench_cfg5_tb14.err
NSB_CONV_TB=2 timeout 300 python bench.py --config 5 --only-headline --no-cpu-baseline > $O/bench_cfg5_tb2.json 2> $O/bench_cfg5_tb2.err
NSB_BENCH_STREAMS=64 NSB_BENCH_R=13 timeout 300 python tools/trace_step.py 2 > $O/trace_cfg5.txt 2>&1
NSB_BENCH_STREAMS=256 NSB_BENCH_R=6 NSB_BENCH_COMPUTE=q8_0 NSB_BENCH_KV=f16 timeout 300 python tools/trace_step.py 2 > $O/trace_cfg3.txt 2>&1
tail -22 $O/trace_cfg5.txt
