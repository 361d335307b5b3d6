#!/bin/bash
O=gpurun_out/c38; mkdir -p $O
run() { local tag=$1; shift
  env "$@" timeout 300 python bench.py --config ${CFG:-3} --only-headline --no-cpu-baseline > $O/$tag.json 2> $O/$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/$tag.json").read().strip().splitlines()[-1]); print("$tag", d["config"]["streams_per_gpu"], round(d["ms_per_step"],3), round(d["value"]), d["breakdown"]["conv_module"]["ms"], d["token_check"]["identical"])
except Exception as e: print("$tag ERR", e)
PY
}
run r2_rule NSB_BENCH_EMULATE_WORLD=2
run r2_tb7 NSB_BENCH_EMULATE_WORLD=2 NSB_CONV_TB=7
run r8_rule NSB_BENCH_EMULATE_WORLD=8
run r4_rule NSB_BENCH_EMULATE_WORLD=4
run c3_rule X=1
CFG=5 run c5_rule X=1
CFG=2 run c2_rule X=1
( timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -k "all_latency or 16bit_and_q8 or large_batch_in_engine or walks_streams or zz_batch or cached_streaming or other_baseline" ) > $O/pytest.log 2>&1; tail -3 $O/pytest.log
