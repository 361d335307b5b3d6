#!/bin/bash
# Q8_0 / Q4_0 through the layer-ahead shadows at every batch size: every quantised-weight test, then bench config 3
O=gpurun_out/c32; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -k "q8 or q4 or Q8 or large_batch or other_baseline or config2_production or zz_batch or serve" ) > $O/pytest.log 2>&1
tail -5 $O/pytest.log
timeout 300 python bench.py --config 3 --only-headline --no-cpu-baseline > $O/bench_cfg3.json 2> $O/bench_cfg3.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c32/bench_cfg3.json').read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3), round(d['value']), round(d['e2e']['value']), d['token_check']['identical'])
PY
