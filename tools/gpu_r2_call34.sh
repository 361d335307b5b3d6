#!/bin/bash
O=gpurun_out/c34; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -k "other_baseline or decode_overlap or walks_streams or large_batch_in_engine or steps_in_flight" ) > $O/pytest.log 2>&1
tail -5 $O/pytest.log
for k in 8 4 2; do
  NSB_BENCH_EMULATE_WORLD=$k timeout 300 python bench.py --config 3 --only-headline --no-cpu-baseline > $O/rank_of_${k}.json 2> $O/rank_of_${k}.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c34/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['config']['streams_per_gpu'], round(d['ms_per_step'],3), round(d['value']), round(d['e2e']['value']), d['latency']['p50'], d['token_check']['identical'])
    except Exception as e: print(f, 'ERR', e)
PY
