#!/bin/bash
# decode overlap at 224 / 448 token rows (one rank of an 8- / 4-GPU strong-scaling run of config 3): automatic rule (off above 128 rows) vs forced
O=gpurun_out/c33; mkdir -p $O
for k in 8 4; do
  NSB_BENCH_EMULATE_WORLD=$k timeout 300 python bench.py --config 3 --only-headline --no-cpu-baseline > $O/rank_of_${k}_auto.json 2> $O/rank_of_${k}_auto.err
  for n in 20 32; do
    NSB_DECODE_OVERLAP=1 NSB_DECODE_CTAS=$n NSB_BENCH_EMULATE_WORLD=$k timeout 300 python bench.py --config 3 --only-headline --no-cpu-baseline > $O/rank_of_${k}_ov$n.json 2> $O/rank_of_${k}_ov$n.err
  done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c33/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['config']['streams_per_gpu'], round(d['ms_per_step'],3), round(d['value']), round(d['e2e']['value']), d['latency']['p50'], d['token_check']['identical'])
    except Exception as e: print(f, 'ERR', e)
PY
