#!/bin/bash
# persistent 256-row pair tiles: correctness, sweep against the one-tile kernel, effect on the config 3 / 5 step; parity of configs 3 / 4 / 5
O=gpurun_out/c13; mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "large_batch_pair_tiles" ) > $O/pytest_gemm.log 2>&1
tail -3 $O/pytest_gemm.log
CFGS="0:0,256:96,208:96,160:96,128:96,112:96,256:96:2,160:96:2,128:96:2,112:96:2,112:97:2" ROWS=1792 timeout 600 python tools/gemm_large.py > $O/gemm_1792.txt 2>&1
CFGS="0:0,256:96,208:96,160:96,128:96,112:96,128:96:2,112:96:2" ROWS=896 R=13 timeout 600 python tools/gemm_large.py > $O/gemm_896.txt 2>&1
cat $O/gemm_1792.txt $O/gemm_896.txt
for c in 3 5; do
  NSB_PAIR256_PERSIST=1 timeout 300 python bench.py --config $c --only-headline --no-cpu-baseline > $O/bench_cfg${c}_persist.json 2> $O/bench_cfg${c}_persist.err
  timeout 300 python bench.py --config $c --only-headline --no-cpu-baseline > $O/bench_cfg${c}.json 2> $O/bench_cfg${c}.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c13/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['ms_per_step'], d['value'], d['e2e']['value'], d['breakdown']['layer_gemm'], d.get('token_check',{}).get('identical'))
    except Exception as e: print(f, 'ERR', e)
PY
( time timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -k "other_baseline or walks_streams" ) > $O/pytest_cfg.log 2>&1
tail -30 $O/pytest_cfg.log
