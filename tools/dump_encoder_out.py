#!/usr/bin/env python3
"""Per-chunk encoder output of a stream in the REFERENCE's own dump format, for front-door parity against a real ggml build
(SURVEY 8f.3): the reference appends `encoder_out` of every chunk to one file (append_dump_tensor / append_dump_array,
src/nemo-stream.cpp:886-958, call site :1009) as a 32-byte header int64 ne[4] (ggml order: ne0 = contiguous dim = 1024, ne1 = frames
per chunk) followed by the f32 payload of chunk after chunk, and scripts/compare_tensors.py diffs two such files.

  python tools/dump_encoder_out.py --impl b200|oracle model.gguf audio.pcm RIGHT_CONTEXT out.bin

--impl b200 runs the CUDA engine through the C ABI (strict-fp32 unless --compute says otherwise) with its debug taps; --impl oracle
runs the CPU checker (test infrastructure). Then, next to a dump made by the reference itself:
  python /path/to/reference/scripts/compare_tensors.py my_bin/ggml_subsampling_output.bin out.bin
"""
import argparse
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def write_dump(path: str, chunks: list) -> None:
    """chunks: list of [T, 1024] float32 arrays (one per processed chunk)."""
    T = chunks[0].shape[0] if chunks else 0
    with open(path, "wb") as f:
        f.write(struct.pack("4q", 1024, T, 1, 1))                      # ggml ne[]: reversed shape
        for c in chunks:
            assert c.shape == (T, 1024)
            f.write(np.ascontiguousarray(c, dtype=np.float32).tobytes())


def encoder_chunks_oracle(model: str, pcm: np.ndarray, R: int) -> list:
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    st = O.Stream(O.Model(model), R, trace=True)
    st.push(pcm)
    return [st.trace_enc(c) for c in range(st.chunks)]


def encoder_chunks_b200(model: str, pcm: np.ndarray, R: int, compute: int) -> list:
    import nsb200
    eng = nsb200.Engine(model, right_context=R, max_streams=1, compute=compute)
    eng.debug_enable(True)
    sid = eng.open_stream()
    eng.push(sid, pcm)
    out = []
    while eng.step() > 0:
        out.append(eng.debug_get("enc", 1))
    eng.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="b200", choices=["b200", "oracle"])
    ap.add_argument("--compute", type=int, default=1, help="nsb_compute for --impl b200 (1 = strict fp32, 0 = from the file)")
    ap.add_argument("model"); ap.add_argument("audio"); ap.add_argument("right_context", type=int); ap.add_argument("out")
    a = ap.parse_args()
    pcm = np.fromfile(a.audio, dtype=np.int16)
    chunks = encoder_chunks_oracle(a.model, pcm, a.right_context) if a.impl == "oracle" else encoder_chunks_b200(a.model, pcm, a.right_context, a.compute)
    write_dump(a.out, chunks)
    print(f"{len(chunks)} chunks x [{chunks[0].shape[0] if chunks else 0}, 1024] -> {a.out}")


if __name__ == "__main__":
    main()
