#!/bin/bash
# decode overlap forced at the large-batch configs: narrow decode (20 / 32 / 48 CTAs) under the next step's encoder
O=gpurun_out/c17; mkdir -p $O
for c in 5 3; do
  for n in 20 32 48; do
    NSB_DECODE_OVERLAP=1 NSB_DECODE_CTAS=$n timeout 300 python bench.py --config $c --only-headline --no-cpu-baseline > $O/bench_cfg${c}_ov$n.json 2> $O/bench_cfg${c}_ov$n.err
  done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c17/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_step'],3), round(d['value']), round(d['e2e']['value']), d.get('latency',{}).get('p50'), d.get('token_check',{}).get('identical'))
    except Exception as e: print(f, 'ERR', e)
PY
