#!/bin/bash
O=gpurun_out/c39; mkdir -p $O
run() { local tag=$1; shift
  env "$@" timeout 300 python bench.py --config 3 --only-headline --no-cpu-baseline > $O/$tag.json 2> $O/$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/$tag.json").read().strip().splitlines()[-1]); print("$tag", d["config"]["streams_per_gpu"], round(d["ms_per_step"],3), round(d["value"]), d["breakdown"]["layer_gemm"]["ms"], {k:v["us"] for k,v in d["roofline"]["per_shape"].items()}, d["token_check"]["identical"])
except Exception as e: print("$tag ERR", e)
PY
}
run r4_base NSB_BENCH_EMULATE_WORLD=4
run r4_pair NSB_BENCH_EMULATE_WORLD=4 NSB_PAIR256_MIN_PAIRS=30
run r8_base NSB_BENCH_EMULATE_WORLD=8
run r8_pair NSB_BENCH_EMULATE_WORLD=8 NSB_PAIR256_MIN_PAIRS=10 NSB_PAIR256_MIN_TILES=2
