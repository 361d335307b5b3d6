#!/usr/bin/env python3
"""Sweep tcgen05 GEMM tile configs on the layer weights (24 layers back to back => weights come from HBM, not L2)."""
import os, sys, itertools
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import nsb200, synth
ROWS = int(os.environ.get("ROWS", 128)); R = 1
path = synth.cached_model("f16", 24, R=R)
eng = nsb200.Engine(path, right_context=R, max_streams=max(64, ROWS // 2), compute=nsb200.COMPUTE_BF16, kv_dtype=nsb200.KV_BF16)
KIND = {0: ("ff1a", 4096, 1024), 1: ("ff1b", 1024, 4096), 2: ("qkv", 3072, 1024), 3: ("out", 1024, 1024), 4: ("pw1", 2048, 1024)}
for kind, (nm, N, K) in KIND.items():
    wbytes = N * K * 2
    for bn, st in ((32, 5), (32, 8), (64, 4), (64, 6), (128, 3), (128, 4), (256, 2), (256, 3)):
        if N % bn: continue
        for splits in (1, 2, 4, 8):
            if splits > 1 and N != 1024: continue
            if (K // 64) % splits or (K // 64) // splits < 2: continue
            for rot in (0, 1):
                try:
                    us = eng.bench_gemm(kind, ROWS, bn, st, splits, rot, 10)
                except Exception as ex:
                    print(nm, bn, st, splits, rot, "ERR", ex); continue
                ctas = (N // bn) * ((ROWS + 127) // 128) * splits
                print(f"{nm:5s} N={N} K={K} rows={ROWS} bn={bn:3d} st={st} splits={splits} rot={rot} ctas={ctas:4d} {us:7.2f} us  {wbytes / us / 1e3:7.1f} GB/s(weights)", flush=True)
eng.close()
