#!/usr/bin/env python3
"""In-graph device timeline of one steady-state engine step (nsb_trace_enable / nsb_trace_fetch: block 0 of every
kernel stamps %globaltimer). Unlike ncu this does not serialise the launches, so it shows what a kernel costs INSIDE
the CUDA graph with programmatic dependent launch: start-to-start pitch, time spent in griddepcontrol.wait, body.
   python tools/trace_step.py [steps] [> gpurun_out/trace.txt]        (same NSB_BENCH_* env as tools/ncu_step.py)
Columns: t0 = block 0 start (us since the step's first record), wait = time until the previous grid had finished,
body = from there to block 0's end, pitch = start-to-start distance to the next kernel.
GEMM rows: pro = prologue (barriers, TMEM alloc), wait = epilogue warps released by the previous grid, mma = accumulator
complete after that, epi = epilogue."""
import collections
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
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    wtype = "q8_0" if COMPUTE == 4 else ("f32" if COMPUTE == 1 else "f16")
    path = synth.cached_model(wtype, N_LAYERS, R=R, profile=os.environ.get("NSB_BENCH_PROFILE", "speech"))
    eng = nsb200.Engine(path, right_context=R, max_streams=STREAMS, compute=COMPUTE, kv_dtype=KV)
    need = 160 * (8 * T * (WARM + 1) - 1) + 256
    base = [synth.synth_pcm(s, need / 16000.0 + 0.01)[:need] for s in range(8)]
    pcm = np.stack([np.roll(base[s % 8], 977 * (s // 8)) for s in range(STREAMS)])
    eng.bench_prepare(pcm, WARM)
    for _ in range(3):
        eng.bench_step()
    eng.trace_enable(4096)
    agg = collections.OrderedDict()
    for it in range(steps):
        ms = eng.bench_step()
        rec = eng.trace_fetch(4096)
        rec.sort(key=lambda r: r[2][0])
        t_first = rec[0][2][0]
        verbose = it == steps - 1
        if verbose:
            print(f"# step {it}: {ms * 1e3:.1f} us by CUDA events, {len(rec)} kernels, first->last block-0 end "
                  f"{(max(max(r[2][:5]) for r in rec) - t_first) / 1e3:.1f} us")
            print(f"{'#':>4s} {'kernel':9s} {'grid':>6s} {'t0':>9s} {'pitch':>7s} | {'pro':>6s} {'wait':>7s} {'mma':>6s} {'body/epi':>8s} {'total':>7s}")
        for i, (tag, grid, t) in enumerate(rec):
            nxt = rec[i + 1][2][0] if i + 1 < len(rec) else None
            pitch = (nxt - t[0]) / 1e3 if nxt else float("nan")
            if tag in ("gemm_tc", "gemm_q8"):
                pro = (t[1] - t[0]) / 1e3; wait = max(0, t[2] - t[1]) / 1e3; mma = max(0, t[3] - max(t[2], t[1])) / 1e3; body = (t[4] - t[3]) / 1e3
                total = (t[4] - t[0]) / 1e3
            elif tag == "attn" and t[3] and t[4]:        # pro = wait-return -> K/V + BD landed; mma = AC + softmax; body = PV + store
                wait = (t[1] - t[0]) / 1e3; pro = (t[3] - t[1]) / 1e3; mma = (t[4] - t[3]) / 1e3; body = (t[2] - t[4]) / 1e3; total = (t[2] - t[0]) / 1e3
            else:
                pro = 0.0; wait = (t[1] - t[0]) / 1e3; mma = 0.0; body = (t[2] - t[1]) / 1e3; total = (t[2] - t[0]) / 1e3
            if tag == "decode" and verbose:
                print(f"#    decode: {t[5] >> 32} rounds ({t[5] & 0xffffffff} with the prediction network): prediction-network phases {t[3] / 1e3:.1f} us, "
                      f"joint + argmax phases {t[4] / 1e3:.1f} us")
            key = f"{tag} g={grid}"
            a = agg.setdefault(key, [0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0])
            a[0] += 1; a[1] += 0 if pitch != pitch else pitch; a[2] += pro; a[3] += wait; a[4] += mma; a[5] += body; a[6] += total
            if verbose:
                print(f"{i:4d} {tag:9s} {grid:6d} {(t[0] - t_first) / 1e3:9.2f} {pitch:7.2f} | {pro:6.2f} {wait:7.2f} {mma:6.2f} {body:8.2f} {total:7.2f}")
    print(f"\n# per kernel class, mean over {steps} step(s): n/step, us/step summed pitch (= its share of the step), mean pro / wait / mma / body / total us")
    tot = sum(a[1] for a in agg.values()) / steps
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        n = a[0]
        print(f"{k:22s} n={n / steps:6.1f} pitch_sum={a[1] / steps:8.1f} us ({100 * a[1] / steps / tot:5.1f} %)  mean: pitch {a[1] / n:6.2f} pro {a[2] / n:5.2f} wait {a[3] / n:6.2f} "
              f"mma {a[4] / n:6.2f} body {a[5] / n:6.2f} total {a[6] / n:6.2f}")
    print(f"# sum of pitches {tot:.1f} us per step")
    eng.close()


if __name__ == "__main__":
    main()
