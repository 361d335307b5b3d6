#!/usr/bin/env python3
"""Wall time of the non-streaming batch path (nsb_transcribe_full) per utterance length: audio seconds per second, 24-layer model."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import nsb200, synth
for wtype, compute in (("f16", nsb200.COMPUTE_BF16), ("f32", nsb200.COMPUTE_F32)):
    eng = nsb200.Engine(synth.cached_model(wtype, 24, R=13), right_context=13, max_streams=1, compute=compute, kv_dtype=nsb200.KV_BF16 if compute != nsb200.COMPUTE_F32 else nsb200.KV_F32)
    for secs in (10.0, 30.0, 60.0, 120.0, 160.0):
        pcm = synth.synth_pcm(5, secs)
        eng.transcribe_full(pcm, want_enc=False)                      # first call: workspace growth
        t0 = time.perf_counter()
        toks, _ = eng.transcribe_full(pcm, want_enc=False)
        dt = time.perf_counter() - t0
        print(f"{wtype} compute={compute} {secs:6.1f} s audio: {dt * 1e3:8.1f} ms  = {secs / dt:8.1f} x real time, {len(toks)} tokens", flush=True)
    if os.environ.get("NSB_PROFILE_BATCH") and compute == nsb200.COMPUTE_BF16:
        pass
    eng.close()
