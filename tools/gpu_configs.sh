#!/bin/bash
# device ms per step for the BASELINE.json configs (steady state); RTFx = streams * 0.08 * (R+1) / (ms / 1000)
run() { echo -n "$1 : "; env $2 timeout 600 python tools/ncu_step.py 6 2>&1 | tail -1; }
run "cfg2 bf16 64x160ms"        "NSB_BENCH_STREAMS=64 NSB_BENCH_R=1 NSB_BENCH_COMPUTE=bf16 NSB_BENCH_KV=bf16"
run "cfg3 q8_0 256x560ms"       "NSB_BENCH_STREAMS=256 NSB_BENCH_R=6 NSB_BENCH_COMPUTE=q8_0 NSB_BENCH_KV=f16"
run "cfg3' f16 256x560ms"       "NSB_BENCH_STREAMS=256 NSB_BENCH_R=6 NSB_BENCH_COMPUTE=f16 NSB_BENCH_KV=f16"
run "cfg4 bf16 128x80ms"        "NSB_BENCH_STREAMS=128 NSB_BENCH_R=0 NSB_BENCH_COMPUTE=bf16 NSB_BENCH_KV=bf16"
run "cfg5 bf16 64x1120ms"       "NSB_BENCH_STREAMS=64 NSB_BENCH_R=13 NSB_BENCH_COMPUTE=bf16 NSB_BENCH_KV=bf16"
run "cfg2 f32 strict 64x160ms"  "NSB_BENCH_STREAMS=64 NSB_BENCH_R=1 NSB_BENCH_COMPUTE=f32 NSB_BENCH_KV=f32"
