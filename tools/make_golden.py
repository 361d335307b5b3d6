#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the REFERENCE's own code (oracle/_ref/libnemo_ref.so = src/preprocessor.cpp +
src/reference/*.cpp compiled where they lie). Run in the build container only (needs /root/reference); the
small fixtures are committed so that the GPU box (no /root/reference) can pin the oracle and the CUDA path."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O  # noqa: E402
import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
CACHED_CASES = [(0, 6.0), (1, 6.6), (6, 7.5), (13, 9.0)]       # (att_right_context, seconds of synth_pcm(11, .)): 74 / 82 / 91 / 112 encoder frames


def main():
    assert O.ref_available(), "oracle/_ref/libnemo_ref.so missing (needs /root/reference)"
    os.makedirs(OUT, exist_ok=True)
    fb = synth.mel_filterbank()
    win = synth.gen_tensor("w", (400,), ("window", 0), 1234, 0, 0)
    # 1. log-mel of the reference preprocessor: synthetic speech-like PCM pushed in CLI-sized reads, and the
    #    reference's own smoke input (0.5 * sin(2 pi 440 t), tests/test_streaming.cpp:745-755)
    pcm = synth.synth_pcm(11, 0.5)
    rp = O.RefPreproc(fb, win)
    mel = np.concatenate([rp.process(pcm[i:i + 2720]) for i in range(0, len(pcm), 2720)])
    sine = synth.sine_pcm(0.25)
    mel_sine = O.RefPreproc(fb, win).process(sine)
    np.savez_compressed(os.path.join(OUT, "mel_ref.npz"), pcm=pcm, mel=mel, sine=sine, mel_sine=mel_sine)
    # 2. model math of src/reference on the 2-layer synthetic model (R=0 calibration)
    rw = O.RefWeights(synth.cached_model("nemo", 2, R=0))
    chunk = np.concatenate([np.zeros((9, 128), np.float32), mel[:8 * 2]])            # first chunk of an R=1 stream: 9 zero frames + 16
    sub = rw.subsampling(chunk)                                                        # [T+2, 1024]
    x = sub[2:]
    layers = []
    for l in range(2):
        x = rw.layer(l, x)
        layers.append(x.copy())
    toks = rw.greedy(x)
    logits0 = rw.joint_logits(x[0], 1024)
    np.savez_compressed(os.path.join(OUT, "model_ref_L2.npz"), chunk=chunk, sub=sub, layer0=layers[0], layer1=layers[1],
                        tokens=toks.astype(np.int32), logits0=logits0)
    # 3. CACHED streaming past the cache roll, every latency mode, free-running on the reference's compiled modules only
    #    (ConvSubsampling::forward per chunk, oracle/ref_shim.cpp:ref_cached_layer_step per layer, GreedyDecoder::decode over all
    #    frames): all tokens + the encoder output of the last chunk (the cache has rolled by then)
    gold = {}
    for R, secs in CACHED_CASES:
        T = 1 + R
        pcm = synth.synth_pcm(11, secs)
        m = O.RefPreproc(fb, win).process(pcm)
        att = [np.zeros((0, 1024), np.float32) for _ in range(2)]
        conv = [np.zeros((0, 1024), np.float32) for _ in range(2)]
        carry, encs = np.zeros((9, 128), np.float32), []
        for c in range(len(m) // (8 * T)):
            ch = np.concatenate([carry, m[8 * T * c:8 * T * (c + 1)]]); carry = ch[-9:]
            x = rw.subsampling(ch)[2:]
            for l in range(2):
                x, att[l], conv[l] = rw.cached_layer_step(l, x, att[l], conv[l])
            encs.append(x)
        assert len(encs) * T > 70 + T
        gold[f"tokens_R{R}"] = rw.greedy(np.concatenate(encs)).astype(np.int32)
        gold[f"enc_last_R{R}"] = encs[-1]
        gold[f"chunks_R{R}"] = np.int32(len(encs))
    np.savez_compressed(os.path.join(OUT, "cached_ref_L2.npz"), **gold)
    # 4. non-streaming batch path (nemo_encode): 3 s utterance, whole-mel subsampling + non-cached layers + greedy from a fresh state
    pcm = synth.synth_pcm(21, 3.0)
    m = O.RefPreproc(fb, win).process(pcm)
    x = rw.subsampling(m)
    for l in range(2):
        x = rw.layer(l, x)
    np.savez_compressed(os.path.join(OUT, "batch_ref_L2.npz"), n_mel=np.int32(len(m)), enc_every4=x[::4], tokens=rw.greedy(x).astype(np.int32))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
