#!/usr/bin/env python3
"""Steady-state engine steps inside a cudaProfilerStart/Stop range, for
   ncu --profile-from-start off ... python tools/ncu_step.py [steps]
Same workload as bench.py (BASELINE.json configs[1] by default; NSB_BENCH_* env overrides)."""
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
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    wtype = "q8_0" if COMPUTE == 4 else ("f32" if COMPUTE == 1 else "f16")
    path = synth.cached_model(wtype, N_LAYERS, R=R, profile=os.environ.get("NSB_BENCH_PROFILE", "speech"))
    eng = nsb200.Engine(path, right_context=R, max_streams=STREAMS, compute=COMPUTE, kv_dtype=KV)
    need = 160 * (8 * T * (WARM + 1) - 1) + 256
    base = [synth.synth_pcm(s, need / 16000.0 + 0.01)[:need] for s in range(8)]
    pcm = np.stack([np.roll(base[s % 8], 977 * (s // 8)) for s in range(STREAMS)])
    eng.bench_prepare(pcm, WARM)
    for _ in range(3):
        eng.bench_step()
    nsb200.lib().nsb_profiler_range(1)
    ms = [eng.bench_step() for _ in range(steps)]
    nsb200.lib().nsb_profiler_range(0)
    print("steps", steps, "ms", ms)
    eng.close()


if __name__ == "__main__":
    main()
