#!/usr/bin/env python3
"""tcgen05 GEMM unit parity vs the oracle's rounding-mirrored matmul (developer tool)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import nsb200, oracle as O, synth  # noqa: E402

def main():
    L = 2
    ok = True
    for wtype, compute, mm in (("f32", 2, O.MM_F16), ("f32", 3, O.MM_BF16), ("f16", 0, O.MM_REF), ("q8_0", 0, O.MM_Q8FAST)):
        path = synth.cached_model(wtype, L, R=0)
        eng = nsb200.Engine(path, right_context=0, max_streams=1, compute=compute)
        om = O.Model(path, mm)
        rng = np.random.default_rng(0)
        P = "encoder.layers.1."
        for name, k in (("feed_forward1.linear1.weight", 1024), ("feed_forward1.linear2.weight", 4096), ("self_attn.linear_out.weight", 1024),
                        ("conv.pointwise_conv1.weight", 1024), ("self_attn.linear_qkv.weight", 1024)):
            for rows in (1, 14, 128, 200, 300):
                x = rng.standard_normal((rows, k)).astype(np.float32)
                y = eng.op_gemm(P + name, x)
                if "qkv" in name:
                    ref = np.concatenate([om.matmul(P + f"self_attn.linear_{c}.weight", x) for c in "qkv"], axis=1)
                else:
                    ref = om.matmul(P + name, x)
                err = float(np.abs(y - ref).max() / np.abs(ref).max())
                flag = "OK " if err < 2e-5 else "BAD"
                if err >= 2e-5: ok = False
                print(f"{flag} wtype={wtype} compute={eng.compute} {name:34s} rows={rows:4d} N={y.shape[1]:5d} relerr={err:.2e}", flush=True)
        eng.close()
    print("ALL OK" if ok else "FAILURES")

if __name__ == "__main__":
    main()
