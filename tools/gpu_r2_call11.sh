#!/bin/bash
# conv module with one block-wide LayerNorm per chunk + attention that walks streams: parity subset, then the bench per config
O=gpurun_out/c11; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "large_batch or walks_streams or all_latency or 16bit_and_q8 or cached_streaming or batch_path or config2" ) > $O/pytest.log 2>&1
tail -4 $O/pytest.log
for c in 3 5 2; do
  timeout 300 python bench.py --config $c --only-headline --no-cpu-baseline > $O/bench_cfg$c.json 2> $O/bench_cfg$c.err
done
NSB_ATT_STREAM=0 timeout 300 python bench.py --config 3 --only-headline --no-cpu-baseline > $O/bench_cfg3_att_old.json 2> $O/bench_cfg3_att_old.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c11/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['ms_per_step'], d['value'], d['e2e']['value'])
    except Exception as e: print(f, 'ERR', e)
PY
