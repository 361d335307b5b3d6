#!/usr/bin/env python3
"""A few engine steps and operator calls that touch every kernel family once, small enough to run under compute-sanitizer
(memcheck / racecheck / synccheck):  compute-sanitizer --tool racecheck python tools/sanitize_run.py
No checker involved: this only has to execute; the sanitizer reports."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import nsb200  # noqa: E402
import synth  # noqa: E402


def stream_case(wtype, compute, kv, R, n, secs, env=None, graph=False):
    for k, v in (env or {}).items():
        os.environ[k] = v
    try:
        eng = nsb200.Engine(synth.cached_model(wtype, 2, R=R), right_context=R, max_streams=n, compute=compute, kv_dtype=kv, cuda_graph=graph)
        base = [synth.synth_pcm(70 + s, secs) for s in range(4)]
        audio = np.stack([base[s % 4] for s in range(n)])
        ids = np.array([eng.open_stream() for _ in range(n)], dtype=np.int32)
        eng.push_batch(ids, audio)
        steps = 0
        while eng.step() > 0:
            steps += 1
        toks = sum(len(eng.pop_tokens(int(i))) for i in ids)
        eng.close()
        print(f"ok stream {wtype} compute={compute} R={R} n={n} env={env}: {steps} steps, {toks} tokens", flush=True)
    finally:
        for k in (env or {}):
            os.environ.pop(k, None)


def gemm_case(wtype, compute, rows, env=None):
    for k, v in (env or {}).items():
        os.environ[k] = v
    try:
        eng = nsb200.Engine(synth.cached_model(wtype, 2, R=0), right_context=0, max_streams=1, compute=compute)
        rng = np.random.default_rng(3)
        for name, k in (("feed_forward1.linear1.weight", 1024), ("feed_forward2.linear2.weight", 4096), ("self_attn.linear_out.weight", 1024)):
            y = eng.op_gemm("encoder.layers.1." + name, rng.standard_normal((rows, k)).astype(np.float32))
            assert np.isfinite(y).all()
        eng.close()
        print(f"ok gemm {wtype} compute={compute} rows={rows} env={env}", flush=True)
    finally:
        for k in (env or {}):
            os.environ.pop(k, None)


if __name__ == "__main__":
    stream_case("f16", nsb200.COMPUTE_BF16, nsb200.KV_BF16, 1, 4, 1.3)                                   # T = 2: pair GEMMs, paired attention, wide decode
    stream_case("f16", nsb200.COMPUTE_BF16, nsb200.KV_BF16, 1, 4, 1.3, {"NSB_DECODE_OVERLAP": "1"})     # narrow decode on its own stream
    stream_case("f16", 0, nsb200.KV_F16, 6, 6, 2.6)                                                      # T = 7: one-item attention, conv module block of 7
    stream_case("f16", 0, nsb200.KV_F16, 6, 6, 2.6, {"NSB_ATT_STREAM": "1"})                             # attention walking its streams
    stream_case("f16", nsb200.COMPUTE_BF16, nsb200.KV_BF16, 13, 80, 3.5)                                 # 1120 rows: pair256 tiles, LayerNorm per warp, conv blocks of 7
    stream_case("q8_0", 0, nsb200.KV_F16, 13, 80, 3.5)                                                   # Q8_0 shadows, dequantisation a layer ahead
    stream_case("f32", nsb200.COMPUTE_F32, nsb200.KV_F32, 0, 3, 1.0)                                     # strict fp32: SIMT GEMM, fp32 attention
    stream_case("q8_0", nsb200.COMPUTE_Q8_0_STRICT, nsb200.KV_F32, 1, 3, 1.0)                           # strict Q8_0
    gemm_case("f32", nsb200.COMPUTE_BF16, 900, {"NSB_PAIR256_PERSIST": "1"})
    gemm_case("q8_0", 0, 900, {"NSB_Q8_PAIR": "1"})
    gemm_case("q4_0", 0, 900, {"NSB_Q8_PAIR": "1"})
    gemm_case("q8_0", 0, 128)
    print("sanitize_run: done")
