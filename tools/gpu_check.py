#!/usr/bin/env python3
"""Exploratory GPU parity run (developer tool; the formal checks live in tests/)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import nsb200  # noqa: E402
import oracle as O  # noqa: E402
import synth  # noqa: E402


def rel(a, b):
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-12))


def main():
    L = int(os.environ.get("L", 2)); R = int(os.environ.get("R", 1)); compute = int(os.environ.get("COMPUTE", 1))
    kv = int(os.environ.get("KV", 0)); nstream = int(os.environ.get("NS", 3)); secs = float(os.environ.get("SECS", 3.0))
    wtype = os.environ.get("WTYPE", "f32")
    mm = {0: (O.MM_Q8FAST if wtype == "q8_0" else O.MM_REF), 1: O.MM_REF, 2: O.MM_F16, 3: O.MM_BF16, 4: O.MM_Q8FAST}[compute]
    path = synth.cached_model(wtype, L, R=R)
    t0 = time.time()
    eng = nsb200.Engine(path, right_context=R, max_streams=nstream, compute=compute, kv_dtype=kv)
    print(f"engine up in {time.time()-t0:.1f}s  layers={eng.n_layers} T={eng.T} compute={eng.compute}", flush=True)

    # ---- mel ----
    pcm = synth.synth_pcm(7, 2.0)
    om = O.Model(path, mm, kv)
    ref_mel = O.Preproc(model=om).process(pcm)
    got = eng.op_logmel(pcm)[0]
    d = np.abs(got.view(np.int32).astype(np.int64) - ref_mel.view(np.int32).astype(np.int64))
    print(f"mel: frames {got.shape[0]} bit-exact {np.mean(d == 0)*100:.2f}%  max ulp {d.max()}  maxabs {np.abs(got-ref_mel).max():.3e}", flush=True)

    # ---- streaming parity ----
    eng.debug_enable(True)
    T = eng.T
    streams = [eng.open_stream() for _ in range(nstream)]
    audio = [synth.synth_pcm(s, secs + 0.37 * s) for s in range(nstream)]
    orc = [O.Stream(om, R, trace=True) for _ in range(nstream)]
    for s in range(nstream):
        orc[s].push(audio[s])
    print("oracle chunks", [o.chunks for o in orc], "tokens", [len(o.tokens()) for o in orc], flush=True)
    # feed the engine in CLI-sized reads and step
    pos = [0] * nstream; done_chunks = [0] * nstream; worst = {}
    read = eng.chunk_samples
    while any(pos[s] < len(audio[s]) for s in range(nstream)):
        for s in range(nstream):
            if pos[s] < len(audio[s]):
                eng.push(streams[s], audio[s][pos[s]:pos[s] + read]); pos[s] += read
        while True:
            ready = [s for s in range(nstream) if eng.ready(streams[s])]
            if not ready:
                break
            B = eng.step()
            assert B == len(ready)
            enc = eng.debug_get("enc", B); sub = eng.debug_get("sub", B); mel = eng.debug_get("mel", B)
            for bi, s in enumerate(ready):
                c = done_chunks[s]
                e_ref = orc[s].trace_enc(c)
                r = rel(enc[bi * T:(bi + 1) * T], e_ref)
                worst["enc"] = max(worst.get("enc", 0), r)
                done_chunks[s] += 1
            if done_chunks[ready[0]] == 1 and ready[0] == 0:
                # deep dive on stream 0 chunk 0 via a second oracle stream replay
                o2 = O.Stream(om, R, trace=True); o2.push(audio[0][: eng.chunk_samples * 2])
                # o2 last_* are for its LAST chunk; only valid if exactly one chunk ran
    for s in range(nstream):
        tg = eng.pop_tokens(streams[s]); to = orc[s].tokens()
        same = len(tg) == len(to) and np.array_equal(tg, to)
        print(f"stream {s}: chunks {eng.chunks(streams[s])}/{orc[s].chunks} tokens {len(tg)}/{len(to)} identical={same}", flush=True)
        if not same:
            n = min(len(tg), len(to)); k = next((i for i in range(n) if tg[i] != to[i]), n)
            print("   first diff at", k, tg[max(0, k-3):k+3], to[max(0, k-3):k+3])
    print("worst rel err:", worst, flush=True)
    st = eng.stats()
    print(f"steps {st.steps} chunks {st.chunks} launches {st.kernel_launches} device_ms {st.device_ms:.2f}", flush=True)

    # ---- per-layer localisation on a fresh single stream, first two chunks ----
    eng2 = nsb200.Engine(path, right_context=R, max_streams=1, compute=compute, kv_dtype=kv); eng2.debug_enable(True)
    s0 = eng2.open_stream(); o = O.Stream(om, R, trace=True)
    a = audio[0]; need = 160 * (8 * T * 1 - 1) + 256
    for c in range(2):
        lo = 0 if c == 0 else need + (c - 1) * 1280 * T; hi = need + c * 1280 * T
        eng2.push(s0, a[lo:hi]); o.push(a[lo:hi]); assert eng2.step() == 1 and o.chunks == c + 1
        print(f" chunk {c}: mel {rel(eng2.debug_get('mel',1)[0], o.last_mel()):.2e} sub {rel(eng2.debug_get('sub',1), o.last_sub()):.2e} " +
              " ".join(f"L{l}:{rel(eng2.debug_get(f'layer.{l}',1), o.last_layer(l)):.1e}" for l in range(eng2.n_layers)), flush=True)
        for which, nm in ((0, "k"), (1, "v"), (2, "conv")):
            print(f"   cache {nm} L0: {rel(eng2.debug_cache(s0, which, 0), o.cache(which, 0)):.2e}", end="")
        lg = eng2.debug_get("logits", 1)
        print(f"   logits evals {lg.shape[0]}: " + (f"{rel(lg[0], o.trace_logits(o.n_evals() - lg.shape[0])):.2e}" if lg.shape[0] else "-"), flush=True)


if __name__ == "__main__":
    main()
