#!/bin/bash
# 1024-stream test, the refreshed default bench line, reference arm
O=gpurun_out/c27; mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "maximum_batch or steps_in_flight" ) > $O/pytest.log 2>&1
tail -4 $O/pytest.log
( time timeout 900 python bench.py ) > $O/bench.json 2> $O/bench.err
tail -c 300 $O/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c27/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['latency']['p50'], d['latency']['p99'], d['roofline']['frac'], d['roofline']['traffic'], d['token_check'])
for k,v in d['configs'].items(): print(k, round(v['ms_per_step'],3), round(v['value']), round(v['e2e']['value']), v['latency_ms']['p50'], v['roofline']['frac'], v['token_check']['identical'])
PY
