#!/usr/bin/env python3
"""Production tile configs only, for A/B runs under env switches (NSB_NO_MC=1, NSB_NO_PDL=1)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import nsb200, synth
ROWS = int(os.environ.get("ROWS", 128))
eng = nsb200.Engine(synth.cached_model("f16", 24, R=1), right_context=1, max_streams=max(64, ROWS // 2), compute=nsb200.COMPUTE_BF16, kv_dtype=nsb200.KV_BF16)
for kind, nm, bn, st, sp in ((0, "ff1a", 32, 5, 1), (2, "qkv", 32, 5, 1), (4, "pw1", 32, 5, 1), (1, "ff1b", 64, 4, 8), (3, "out", 32, 5, 4), (0, "ff1a", 64, 4, 1), (0, "ff1a", 128, 4, 1)):
    us = eng.bench_gemm(kind, ROWS, bn, st, sp, 1, 20)
    print(f"{nm:5s} rows={ROWS} bn={bn} st={st} splits={sp}: {us:6.2f} us", flush=True)
eng.close()
