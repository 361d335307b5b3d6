#!/usr/bin/env python3
"""Per-evaluation logit distance engine vs oracle for ONE stream (row 0 owns the logits tap). Developer diagnostic."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import nsb200, synth
import oracle as O

def main():
    seed = int(os.environ.get("SEED", 51)); secs = float(os.environ.get("SECS", 2.3)); R = int(os.environ.get("R", 1))
    wtype = os.environ.get("WTYPE", "f16"); compute = int(os.environ.get("COMPUTE", 0)); kv = int(os.environ.get("KV", 1))
    mm = {"f16": O.MM_REF, "f32": O.MM_F16}[wtype]; okv = {0: O.KV_F32, 1: O.KV_F16, 2: O.KV_BF16}[kv]
    path = synth.cached_model(wtype, 2, R=R)
    eng = nsb200.Engine(path, right_context=R, max_streams=1, compute=compute, kv_dtype=kv); eng.debug_enable(True)
    om = O.Model(path, mm, okv)
    pcm = synth.synth_pcm(seed, secs)
    o = O.Stream(om, R, trace=True); o.push(pcm)
    sid = eng.open_stream(); T = R + 1
    read = eng.chunk_samples; pos = 0; ev0 = 0; chunk = 0; worst_l = 0.0; worst_e = 0.0; flips = 0
    while pos < len(pcm):
        eng.push(sid, pcm[pos:pos + read]); pos += read
        while eng.ready(sid):
            assert eng.step() == 1
            enc = eng.debug_get("enc", 1); er = float(np.abs(enc - o.trace_enc(chunk)).max() / (np.abs(o.trace_enc(chunk)).max() + 1e-12))
            lg = eng.debug_get("logits", 1)
            msg = []
            for i in range(lg.shape[0]):
                if ev0 + i >= o.n_evals(): break
                ol = o.trace_logits(ev0 + i); d = float(np.abs(lg[i] - ol).max()); s = np.sort(ol)
                worst_l = max(worst_l, d)
                flip = int(np.argmax(lg[i])) != int(np.argmax(ol)); flips += flip
                msg.append(f"{d:.3f}/{s[-1]-s[-2]:.2f}{'!' if flip else ''}")
                if flip: break
            worst_e = max(worst_e, er)
            print(f"chunk {chunk}: enc rel {er:.2e}  evals {lg.shape[0]}  |dlogit|/oracle-gap: {' '.join(msg[:14])}")
            ev0 += lg.shape[0]; chunk += 1
            if flips: break
        if flips: break
    print(f"worst enc rel {worst_e:.2e} worst |dlogit| {worst_l:.3f} first flip seen: {bool(flips)}  tokens engine {len(eng.pop_tokens(sid))} oracle {len(o.tokens())}")

main()
