#!/usr/bin/env python3
"""Layer GEMM shapes at a large batch (ROWS token rows, default 1792 = 256 streams x 560 ms): us and TFLOP/s per tile config.
bn = 0 / stages = 0 is the engine's own choice. Timed as in bench.py (24 layers x iters launches, one CUDA graph per pass)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import nsb200, synth
ROWS = int(os.environ.get("ROWS", 1792)); R = int(os.environ.get('R', 6)); T = R + 1
CFGS = [tuple((list(map(int, c.split(":"))) + [1])[:3]) for c in os.environ.get("CFGS", "0:0,256:2,256:3,128:3,128:4,128:6,64:4").split(",")]   # bn:stages[:splits]; stages 97 = 256-row CTA-pair tile, 96 = its persistent variant; 0:0:-1 = the step's own launch
Q8 = os.environ.get("COMPUTE", "bf16") == "q8_0"          # Q8_0 weights: stages 95 = fused dequantisation on the 256-row pair tiles; 0:0 = the engine's own path
eng = (nsb200.Engine(synth.cached_model("q8_0", 24, R=R), right_context=R, max_streams=(ROWS + T - 1) // T, compute=0, kv_dtype=nsb200.KV_F16) if Q8 else
       nsb200.Engine(synth.cached_model("f16", 24, R=R), right_context=R, max_streams=(ROWS + T - 1) // T, compute=nsb200.COMPUTE_BF16, kv_dtype=nsb200.KV_BF16))
KIND = {0: ("ff_up", 4096, 1024), 1: ("ff_down", 1024, 4096), 2: ("qkv", 3072, 1024), 3: ("out", 1024, 1024), 4: ("pw1", 2048, 1024)}
for kind, (nm, N, K) in KIND.items():
    for bn, st, sp in CFGS:
        if bn and st not in (95, 96, 97) and N % bn: continue
        if sp > 1 and N != 1024: continue
        try:
            us = eng.bench_gemm(kind, ROWS, bn, st, sp, 1, 10)
        except Exception as ex:
            print(nm, bn, st, "ERR", ex); continue
        print(f"{nm:7s} N={N} K={K} rows={ROWS} bn={bn:3d} st={st} splits={sp}: {us:7.2f} us  {2.0 * ROWS * N * K / us / 1e6:7.1f} TFLOP/s", flush=True)
eng.close()
