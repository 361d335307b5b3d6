#!/bin/bash
O=gpurun_out/c3; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -k "batch_path or strict_q8 or q8 or large_batch or config2 or overlap or loader or graph_replay" ) > $O/pytest.log 2>&1
tail -4 $O/pytest.log
( time timeout 900 python bench.py ) > $O/bench.json 2> $O/bench.err
tail -c 300 $O/bench.err
NSB_Q8_SHADOW=0 timeout 300 python bench.py --config 3 --only-headline --no-cpu-baseline > $O/bench_cfg3_noshadow.json 2> $O/bench_cfg3_noshadow.err
NSB_BENCH_STREAMS=256 NSB_BENCH_R=6 NSB_BENCH_COMPUTE=q8_0 NSB_BENCH_KV=f16 timeout 300 python tools/trace_step.py 2 > $O/trace_cfg3.txt 2>&1
tail -25 $O/trace_cfg3.txt
