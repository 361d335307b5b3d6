#!/bin/bash
# three steps in flight: API tests, production-path parity, bench config 2 / 4 with 3 and 2 steps in flight in the e2e leg
O=gpurun_out/c25; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "in_flight or split_step or config2_production or graph_replay or decode_overlap or serve or api_edge or batch_push" ) > $O/pytest.log 2>&1
tail -4 $O/pytest.log
for c in 2 4; do
  timeout 300 python bench.py --config $c --only-headline --no-cpu-baseline > $O/bench_cfg${c}_if3.json 2> $O/bench_cfg${c}_if3.err
  NSB_BENCH_IN_FLIGHT=2 timeout 300 python bench.py --config $c --only-headline --no-cpu-baseline > $O/bench_cfg${c}_if2.json 2> $O/bench_cfg${c}_if2.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c25/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_step'],3), round(d['value']), round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3), d['latency']['p50'], d.get('token_check',{}).get('identical'))
    except Exception as e: print(f, 'ERR', e)
PY
