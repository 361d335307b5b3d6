#!/bin/bash
# PCM upload on its own stream: API / in-flight / production-path tests, then e2e of configs 3, 5, 2
O=gpurun_out/c43; mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "in_flight or split_step or randomised or api_edge or batch_push or graph_replay or decode_overlap or serve or two_gpus or (other_baseline and config3_q8)" ) > $O/pytest.log 2>&1
tail -4 $O/pytest.log
for c in 3 5 2; do
  timeout 300 python bench.py --config $c --only-headline --no-cpu-baseline > $O/cfg$c.json 2> $O/cfg$c.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/cfg$c.json").read().strip().splitlines()[-1]); print($c, round(d["ms_per_step"],3), round(d["value"]), round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],3), d["latency"]["p50"], d["token_check"]["identical"])
except Exception as e: print($c, "ERR", e)
PY
done
