#!/bin/bash
O=gpurun_out/c23; mkdir -p $O
( timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "large_batch_pair_tiles and (q8 or q4)" ) > $O/pytest_gemm.log 2>&1
tail -2 $O/pytest_gemm.log
COMPUTE=q8_0 CFGS="0:0,208:95,160:95,112:95,112:95:2,160:95:2" ROWS=1792 timeout 600 python tools/gemm_large.py > $O/gemm_q8_1792.txt 2>&1
cat $O/gemm_q8_1792.txt
timeout 300 python bench.py --config 3 --only-headline --no-cpu-baseline > $O/bench_cfg3_q8pair.json 2> $O/bench_cfg3_q8pair.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c23/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['ms_per_step'], d['value'], d['e2e']['value'], d['breakdown']['layer_gemm'], d.get('token_check',{}).get('identical'))
    except Exception as e: print(f, 'ERR', e)
PY
