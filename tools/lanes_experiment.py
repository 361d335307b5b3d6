#!/usr/bin/env python3
"""Experiment: does the GPU take two half-batches concurrently better than one full batch?
One engine with S streams vs L engines with S/L streams each, every engine on its own CUDA stream and host thread
(the step of a small batch is a chain of latency-bound kernels; two chains interleave on the SMs).
   python tools/lanes_experiment.py [steps]      env: NSB_BENCH_STREAMS (64), NSB_BENCH_R (1), LANES ("1,2,4")"""
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import nsb200  # noqa: E402
import synth  # noqa: E402

STREAMS = int(os.environ.get("NSB_BENCH_STREAMS", 64))
R = int(os.environ.get("NSB_BENCH_R", 1))
T = 1 + R
WARM = 70 // T + 3
CHUNKS = 8


def make(n_streams, first):
    path = synth.cached_model("f16", 24, R=R, profile="speech")
    eng = nsb200.Engine(path, right_context=R, max_streams=n_streams, compute=nsb200.COMPUTE_BF16, kv_dtype=nsb200.KV_BF16)
    need = 160 * (8 * T * (WARM + CHUNKS) - 1) + 256
    base = [synth.synth_pcm(1000 + s, need / 16000.0 + 0.01)[:need] for s in range(8)]
    pcm = np.stack([np.roll(base[s % 8], 977 * (s // 8)) for s in range(first, first + n_streams)])
    eng.bench_prepare(pcm, WARM)
    for _ in range(CHUNKS):
        eng.bench_step()
    return eng


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    for lanes in [int(x) for x in os.environ.get("LANES", "1,2,4").split(",")]:
        per = STREAMS // lanes
        engs = [make(per, i * per) for i in range(lanes)]
        res = [None] * lanes
        bar = threading.Barrier(lanes + 1)

        def work(i):
            bar.wait()
            res[i] = engs[i].bench_steps(steps)[0]

        th = [threading.Thread(target=work, args=(i,)) for i in range(lanes)]
        for t in th:
            t.start()
        bar.wait()
        t0 = time.perf_counter()
        for t in th:
            t.join()
        wall = time.perf_counter() - t0
        audio = STREAMS * 0.08 * T * steps
        print(f"lanes={lanes} streams/lane={per} steps={steps}: wall {1e3 * wall / steps:.3f} ms per step-of-all-lanes, "
              f"device ms per lane {[round(r / steps, 3) for r in res]}, RTFx {audio / wall:.0f}", flush=True)
        for e in engs:
            e.close()


if __name__ == "__main__":
    main()
