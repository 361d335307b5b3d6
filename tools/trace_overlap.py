#!/usr/bin/env python3
"""Back-to-back steps (nsb_bench_steps) under the in-graph device trace: where the decode of step i sits relative to the encoder
kernels of step i + 1 (decode overlap), and what a step costs start to start.
   python tools/trace_overlap.py [steps]          (NSB_BENCH_* env as tools/trace_step.py; NSB_DECODE_OVERLAP=0 for the inline decode)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import nsb200  # noqa: E402
import synth  # noqa: E402

N_LAYERS = int(os.environ.get("NSB_BENCH_LAYERS", 24))
STREAMS = int(os.environ.get("NSB_BENCH_STREAMS", 64))
R = int(os.environ.get("NSB_BENCH_R", 1))
COMPUTE = {"f32": 1, "f16": 2, "bf16": 3, "q8_0": 4}[os.environ.get("NSB_BENCH_COMPUTE", "bf16")]
KV = {"f32": 0, "f16": 1, "bf16": 2}[os.environ.get("NSB_BENCH_KV", "bf16")]
T = 1 + R
WARM = 70 // T + 3


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    wtype = "q8_0" if COMPUTE == 4 else ("f32" if COMPUTE == 1 else "f16")
    path = synth.cached_model(wtype, N_LAYERS, R=R, profile=os.environ.get("NSB_BENCH_PROFILE", "speech"))
    eng = nsb200.Engine(path, right_context=R, max_streams=STREAMS, compute=COMPUTE, kv_dtype=KV)
    need = 160 * (8 * T * (WARM + 9) - 1) + 256
    base = [synth.synth_pcm(s, need / 16000.0 + 0.01)[:need] for s in range(8)]
    pcm = np.stack([np.roll(base[s % 8], 977 * (s // 8)) for s in range(STREAMS)])
    eng.bench_prepare(pcm, WARM)
    for _ in range(16):
        eng.bench_step()
    eng.trace_enable(8192)
    tot, per = eng.bench_steps(steps)
    rec = eng.trace_fetch(8192)
    rec.sort(key=lambda r: r[2][0])
    t0 = rec[0][2][0]
    print(f"# {steps} steps back to back: {tot * 1e3:.1f} us by CUDA events ({tot * 1e3 / steps:.1f} us per step), per step {[round(p * 1e3, 1) for p in per]}")
    # step boundaries = log-mel kernels; decode kernels reported with their span
    mel = [r for r in rec if r[0] == "logmel"]
    dec = [r for r in rec if r[0] == "decode"]
    for i, m in enumerate(mel):
        end_enc = None
        nxt = mel[i + 1][2][0] if i + 1 < len(mel) else None
        kern = [r for r in rec if r[0] != "decode" and r[2][0] >= m[2][0] and (nxt is None or r[2][0] < nxt)]
        end_enc = max(max(r[2][:5]) for r in kern)
        print(f"step {i}: encoder {len(kern)} kernels, start {(m[2][0] - t0) / 1e3:9.1f} us, last block-0 end {(end_enc - t0) / 1e3:9.1f} us "
              f"({(end_enc - m[2][0]) / 1e3:7.1f} us)")
    for i, d in enumerate(dec):
        t = d[2]
        print(f"decode {i}: grid {d[1]:4d}, start {(t[0] - t0) / 1e3:9.1f} us, end {(t[2] - t0) / 1e3:9.1f} us ({(t[2] - t[0]) / 1e3:7.1f} us), "
              f"{t[5] >> 32} rounds ({t[5] & 0xffffffff} with the prediction network): prediction {t[3] / 1e3:.1f} us, joint {t[4] / 1e3:.1f} us")
    eng.close()


if __name__ == "__main__":
    main()
