#!/bin/bash
# bench under launch-mode variants (device ms/step only)
OUT=gpurun_out/${1:-var}; mkdir -p $OUT
for v in "default" "NSB_NO_PDL=1" "NSB_BENCH_GRAPH=0" "NSB_NO_PDL=1 NSB_BENCH_GRAPH=0"; do
  if [ "$v" = "default" ]; then E=""; else E="$v"; fi
  env $E python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $OUT/b.json 2>$OUT/b.err
  python -c "import json;d=json.load(open('$OUT/b.json'));print('$v', 'ms/step', round(d['ms_per_step'],3), 'rtfx', round(d['value']), 'e2e', round(d['e2e']['value']))"
done
