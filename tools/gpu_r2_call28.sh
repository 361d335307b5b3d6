#!/bin/bash
# config 3 strong scaling, one rank's share of a 2 / 4 / 8-GPU run on one GPU (128 / 64 / 32 streams x 560 ms, Q8_0)
O=gpurun_out/c28; mkdir -p $O
for k in 2 4 8; do
  NSB_BENCH_EMULATE_WORLD=$k timeout 300 python bench.py --config 3 --only-headline --no-cpu-baseline > $O/bench_cfg3_rank_of_$k.json 2> $O/bench_cfg3_rank_of_$k.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c28/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['config']['streams_per_gpu'], round(d['ms_per_step'],3), round(d['value']), round(d['e2e']['value']), d['latency']['p50'], d['token_check']['identical'], d['roofline']['frac'], d['roofline']['bound'])
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-500:])
PY
