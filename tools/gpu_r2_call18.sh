#!/bin/bash
# Q8_0 / Q4_0 dequantisation fused into the 256-row CTA-pair tiles: GEMM parity, tile sweep, config 3 step against the shadow path
O=gpurun_out/c18; mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "large_batch_pair_tiles" ) > $O/pytest_gemm.log 2>&1
tail -3 $O/pytest_gemm.log
COMPUTE=q8_0 CFGS="0:0,208:95,160:95,112:95,112:95:2,160:95:2,208:95:2" ROWS=1792 timeout 600 python tools/gemm_large.py > $O/gemm_q8_1792.txt 2>&1
cat $O/gemm_q8_1792.txt
timeout 300 python bench.py --config 3 --only-headline --no-cpu-baseline > $O/bench_cfg3_q8pair.json 2> $O/bench_cfg3_q8pair.err
NSB_Q8_PAIR=0 timeout 300 python bench.py --config 3 --only-headline --no-cpu-baseline > $O/bench_cfg3_shadow.json 2> $O/bench_cfg3_shadow.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c18/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['ms_per_step'], d['value'], d['e2e']['value'], d['breakdown']['layer_gemm'], d.get('token_check',{}).get('identical'))
    except Exception as e: print(f, 'ERR', e)
PY
( time timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -k "other_baseline or large_batch_in_engine or 16bit_and_q8 or q8_fast" ) > $O/pytest_cfg.log 2>&1
tail -8 $O/pytest_cfg.log
