#!/bin/bash
# conv module: frames per block at small batches (more CTAs against recomputed window rows)
O=gpurun_out/c37; mkdir -p $O
run() { # tag env...
  local tag=$1; shift
  env "$@" timeout 300 python bench.py --config ${CFG:-3} --only-headline --no-cpu-baseline > $O/$tag.json 2> $O/$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/$tag.json").read().strip().splitlines()[-1]); print("$tag", d["config"]["streams_per_gpu"], round(d["ms_per_step"],3), round(d["value"]), d["breakdown"]["conv_module"]["ms"], d["token_check"]["identical"])
except Exception as e: print("$tag ERR", e)
PY
}
for tb in 7 4 3 2; do run r8_tb$tb NSB_BENCH_EMULATE_WORLD=8 NSB_CONV_TB=$tb; done
for tb in 7 4 3; do run r4_tb$tb NSB_BENCH_EMULATE_WORLD=4 NSB_CONV_TB=$tb; done
CFG=5; for tb in 7 5 4; do run c5_tb$tb NSB_CONV_TB=$tb; done
