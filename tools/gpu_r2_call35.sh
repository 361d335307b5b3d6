#!/bin/bash
# closing check at HEAD: full GPU suite, smoke, default bench line (+ reference arm)
O=gpurun_out/c44; mkdir -p $O
( time timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=5 ) > $O/pytest.log 2>&1
tail -12 $O/pytest.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/smoke.log 2>&1; echo "smoke rc=$?"; head -2 $O/smoke.log
( time timeout 900 python bench.py ) > $O/bench.json 2> $O/bench.err
tail -c 200 $O/bench.err
true
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c44/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['latency']['p50'], d['latency']['p99'], d['roofline']['frac'], d['cpu_baseline']['value'], d['token_check']['identical'])
for k,v in d['configs'].items(): print(k, round(v['ms_per_step'],3), round(v['value']), round(v['e2e']['value']), v['latency_ms']['p50'], v['roofline']['frac'], v['token_check']['identical'])

PY
