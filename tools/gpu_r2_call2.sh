#!/bin/bash
# round 2, GPU call 2: full GPU suite (loader rewrite, decode overlap, batch path changes), bench with / without decode overlap,
# in-graph traces (single step per config, back-to-back steps), ncu launch list of the headline config
O=gpurun_out/c2; mkdir -p $O
( time timeout 1800 python -m pytest tests -m gpu -q --durations=12 -p no:cacheprovider ) > $O/pytest.log 2>&1
tail -4 $O/pytest.log
( time timeout 900 python bench.py ) > $O/bench.json 2> $O/bench.err
tail -c 300 $O/bench.err
NSB_DECODE_OVERLAP=0 timeout 300 python bench.py --only-headline --no-cpu-baseline > $O/bench_no_overlap.json 2> $O/bench_no_overlap.err
timeout 300 python tools/trace_overlap.py 4 > $O/trace_overlap.txt 2>&1
NSB_DECODE_OVERLAP=0 timeout 300 python tools/trace_overlap.py 4 > $O/trace_no_overlap.txt 2>&1
timeout 300 python tools/trace_step.py 2 > $O/trace_cfg2.txt 2>&1
NSB_BENCH_STREAMS=256 NSB_BENCH_R=6 NSB_BENCH_COMPUTE=q8_0 NSB_BENCH_KV=f16 timeout 300 python tools/trace_step.py 2 > $O/trace_cfg3.txt 2>&1
NSB_BENCH_STREAMS=64 NSB_BENCH_R=13 timeout 300 python tools/trace_step.py 2 > $O/trace_cfg5.txt 2>&1
timeout 200 python tools/ncu_step.py 2 > $O/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches.csv python tools/ncu_step.py 2 > $O/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
for f in trace_overlap trace_no_overlap; do head -12 $O/$f.txt; done
