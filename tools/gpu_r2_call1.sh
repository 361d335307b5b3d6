#!/bin/bash
# round 2, GPU call 1: the whole GPU suite (incl. the new production-path parity tests), the bench line with every config,
# 3xTF32 vs single-pass tf32 A/B (error + step time), smoke
O=gpurun_out/c1; mkdir -p $O
nvidia-smi -L > $O/smi.txt; nproc >> $O/smi.txt
( time timeout 1800 python -m pytest tests -m gpu -q --durations=25 -p no:cacheprovider ) > $O/pytest.log 2>&1
tail -5 $O/pytest.log
( time timeout 900 python bench.py ) > $O/bench.json 2> $O/bench.err
tail -c 600 $O/bench.err
NSB_STEM_TF32=1 timeout 300 python bench.py --only-headline --no-cpu-baseline > $O/bench_tf32x1.json 2> $O/bench_tf32x1.err
NSB_STEM_TF32=1 timeout 600 python -m pytest tests/test_gpu_production_path.py -k "encoder_error" -q -p no:cacheprovider > $O/pytest_tf32x1.log 2>&1
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1
tail -3 $O/smoke.log
