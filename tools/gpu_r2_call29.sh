#!/bin/bash
# Q8_0 at 224 / 448 token rows: fused per-m-tile dequantisation (default below 512 rows) against the layer-ahead shadows
O=gpurun_out/c29; mkdir -p $O
for k in 4 8; do
  for thr in 512 192; do
    NSB_Q8_PREDEQUANT_ROWS=$thr NSB_BENCH_EMULATE_WORLD=$k timeout 300 python bench.py --config 3 --only-headline --no-cpu-baseline > $O/bench_rank_of_${k}_thr$thr.json 2> $O/bench_rank_of_${k}_thr$thr.err
  done
done
# the same shapes in bf16 for reference
for s in 64 32; do NSB_BENCH_STREAMS=$s NSB_BENCH_R=6 timeout 300 python bench.py --config 2 --only-headline --no-cpu-baseline > $O/bench_bf16_$s.json 2> $O/bench_bf16_$s.err; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c29/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['config']['streams_per_gpu'], round(d['ms_per_step'],3), round(d['value']), round(d['e2e']['value']), d['token_check']['identical'], d['breakdown']['layer_gemm']['ms'])
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
PY
