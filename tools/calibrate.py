#!/usr/bin/env python3
"""Compute the mean encoder output mu of a synthetic model (seed, n_layers) with the CPU oracle
and store it under tools/calib/ so that tools/synth.py can fold -W_enc@mu into joint.enc.bias.
Run once per (seed, n_layers); the result is committed (4 KB). Test tooling only."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", "oracle"))
import oracle as O  # noqa: E402
import synth  # noqa: E402


def main():
    n_layers = int(sys.argv[1]); seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1234
    out = os.path.join(HERE, "calib", f"mu_s{seed}_L{n_layers}.npy")
    if os.path.exists(out):
        os.remove(out)
    path = f"/tmp/calib_s{seed}_L{n_layers}.gguf"
    synth.write_gguf(path, n_layers, "f32", seed)
    m = O.Model(path)
    rows = []
    for stream in (100, 101, 102):
        s = O.Stream(m, 13, trace=True)
        s.push(synth.synth_pcm(stream, 3.5))
        rows += [s.trace_enc(c) for c in range(s.chunks)]
    E = np.concatenate(rows)
    np.save(out, E.mean(axis=0).astype(np.float32))
    os.remove(path)
    print(out, E.shape, "mean |mu|", float(np.abs(E.mean(axis=0)).mean()))


if __name__ == "__main__":
    main()
