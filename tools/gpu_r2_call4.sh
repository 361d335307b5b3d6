#!/bin/bash
O=gpurun_out/c4; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "not config2 and not encoder_error and not strict_q8" ) > $O/pytest.log 2>&1
tail -4 $O/pytest.log
timeout 300 python tools/trace_overlap.py 4 > $O/trace_inline.txt 2>&1
NSB_DECODE_OVERLAP=1 NSB_DECODE_CTAS=20 timeout 300 python tools/trace_overlap.py 4 > $O/trace_overlap20.txt 2>&1
NSB_DECODE_OVERLAP=1 NSB_DECODE_CTAS=16 timeout 300 python tools/trace_overlap.py 4 > $O/trace_overlap16.txt 2>&1
NSB_BENCH_STREAMS=128 NSB_BENCH_R=0 NSB_DECODE_OVERLAP=1 NSB_DECODE_CTAS=20 timeout 300 python tools/trace_overlap.py 4 > $O/trace_overlap20_cfg4.txt 2>&1
NSB_BENCH_STREAMS=128 NSB_BENCH_R=0 timeout 300 python tools/trace_overlap.py 4 > $O/trace_inline_cfg4.txt 2>&1
NSB_BENCH_STREAMS=256 NSB_BENCH_R=6 NSB_BENCH_COMPUTE=q8_0 NSB_BENCH_KV=f16 timeout 300 python tools/trace_overlap.py 3 > $O/trace_cfg3.txt 2>&1
NSB_DECODE_PIPE_N=100000 NSB_BENCH_STREAMS=256 NSB_BENCH_R=6 NSB_BENCH_COMPUTE=q8_0 NSB_BENCH_KV=f16 timeout 300 python tools/trace_overlap.py 3 > $O/trace_cfg3_oldphase.txt 2>&1
NSB_BENCH_STREAMS=64 NSB_BENCH_R=13 timeout 300 python tools/trace_overlap.py 3 > $O/trace_cfg5.txt 2>&1
NSB_DECODE_PIPE_N=17 NSB_BENCH_STREAMS=64 NSB_BENCH_R=13 timeout 300 python tools/trace_overlap.py 3 > $O/trace_cfg5_pipe17.txt 2>&1
NSB_DECODE_PIPE_N=17 timeout 300 python tools/trace_overlap.py 4 > $O/trace_inline_pipe17.txt 2>&1
NSB_DECODE_OVERLAP=1 NSB_DECODE_CTAS=20 timeout 300 python bench.py --only-headline --no-cpu-baseline > $O/bench_overlap20.json 2> $O/bench_overlap20.err
for f in $O/trace_*.txt; do echo "== $f"; grep -v "^step" $f | head -8; done
