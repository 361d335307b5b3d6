#!/bin/bash
O=gpurun_out/c10; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "q4 or q8 or large_batch or batch_path or cached_streaming or f32_all_latency" ) > $O/pytest.log 2>&1
tail -4 $O/pytest.log
( time timeout 900 python bench.py --no-cpu-baseline ) > $O/bench.json 2> $O/bench.err
tail -c 300 $O/bench.err
NSB_CONV_TB=4 timeout 300 python bench.py --config 3 --only-headline --no-cpu-baseline > $O/bench_cfg3_tb4.json 2> $O/bench_cfg3_tb4.err
