#!/bin/bash
# compute-sanitizer over every kernel family (memcheck, then racecheck + synccheck on the same small run)
O=gpurun_out/c24; mkdir -p $O
timeout 300 python tools/sanitize_run.py > $O/plain.log 2>&1; echo "plain rc=$?"; tail -3 $O/plain.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 30 python tools/sanitize_run.py > $O/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 $O/memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 --print-limit 30 python tools/sanitize_run.py > $O/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 $O/racecheck.log
timeout 1500 compute-sanitizer --tool synccheck --error-exitcode 9 --print-limit 30 python tools/sanitize_run.py > $O/synccheck.log 2>&1; echo "synccheck rc=$?"; tail -4 $O/synccheck.log
timeout 300 python bench.py --config 3 --only-headline --no-cpu-baseline > $O/bench_cfg3.json 2> $O/bench_cfg3.err; cut -c1-1200 $O/bench_cfg3.json | tail -c 700
