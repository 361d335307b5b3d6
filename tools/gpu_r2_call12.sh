#!/bin/bash
# conv module (rolled loop, one barrier) timing per config, then ncu --set full with source: narrow decode (config 2), and at config 3
# the attention kernels (one item per CTA / walking streams), the conv module and the full-width decode
O=gpurun_out/c12; mkdir -p $O
for c in 2 5 3; do
  timeout 300 python bench.py --config $c --only-headline --no-cpu-baseline > $O/bench_cfg$c.json 2> $O/bench_cfg$c.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c12/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['ms_per_step'], d['value'], d['e2e']['value'], d['breakdown']['conv_module'], d['breakdown']['attention'])
    except Exception as e: print(f, 'ERR', e)
PY
export KEEP_REP=1
NSB_DECODE_OVERLAP=1 bash tools/gpu_ncu_one.sh rnnt_decode 1 c12/dec_narrow_cfg2 > $O/narrow.log 2>&1
export NSB_BENCH_STREAMS=256 NSB_BENCH_R=6 NSB_BENCH_COMPUTE=q8_0 NSB_BENCH_KV=f16
bash tools/gpu_ncu_one.sh "attention_mma|conv_module" 2 c12/att_conv_cfg3 > $O/att.log 2>&1
NSB_ATT_STREAM=1 bash tools/gpu_ncu_one.sh "attention_mma" 1 c12/att_stream_cfg3 > $O/att_stream.log 2>&1
bash tools/gpu_ncu_one.sh rnnt_decode 1 c12/dec_wide_cfg3 > $O/wide.log 2>&1
for r in dec_narrow_cfg2 dec_wide_cfg3 att_conv_cfg3 att_stream_cfg3; do
  ncu -i $O/$r.ncu-rep --page source --csv > $O/${r}_source.csv 2>/dev/null
done
ls -la $O; tail -2 $O/*.log
