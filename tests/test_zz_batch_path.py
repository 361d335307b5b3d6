"""Non-streaming batch path (nsb_transcribe_full) on the GPU against its checker: strict fp32 against the CPU checker and
the fixture produced by the reference's compiled modules (tokens identical, token frames identical, encoder <= 1e-4), fp16, an engine
with ONE stream slot taking utterances of any length (the batch workspace is its own), the reference's src/transcribe.cpp drop-in, and
the full-depth model on 30 s of audio. Runs in a SUBPROCESS, last in the suite (a faulting kernel cannot take the CUDA context of the
parity suite with it). The checker side is pinned on CPU (tests/test_oracle.py::test_batch_path_matches_compiled_reference)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import os, sys
ROOT = sys.argv[1]
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np
import nsb200, synth, oracle as O

path = synth.cached_model("f32", 2, R=0)
pcm = synth.synth_pcm(21, 3.0)
om = O.Model(path)
mel = O.Preproc(model=om).process(pcm)
enc_o, toks_o, frames_o = om.transcribe_full(mel)
g = np.load(os.path.join(ROOT, "tests", "golden", "batch_ref_L2.npz"))

eng = nsb200.Engine(path, right_context=13, max_streams=1, compute=nsb200.COMPUTE_F32)      # one slot: the batch workspace is its own
toks, enc, frames = eng.transcribe_full(pcm, want_frames=True)
assert enc.shape == enc_o.shape, (enc.shape, enc_o.shape)
rel = float(np.abs(enc - enc_o).max() / np.abs(enc_o).max())
print("strict fp32: encoder rel err vs oracle", rel, "tokens", len(toks), len(toks_o))
assert rel < 1e-4, rel
assert float(np.abs(enc[::4] - g["enc_every4"]).max() / np.abs(g["enc_every4"]).max()) < 1e-4
assert np.array_equal(toks, toks_o) and np.array_equal(toks, g["tokens"])
assert np.array_equal(frames, frames_o), (frames, frames_o)                                   # timed_token::frame_idx (nemo-ggml.cpp:1240)
# the slot and the step workspace were only borrowed: a streaming stream on the same engine still matches the oracle
sid = eng.open_stream(); eng.push(sid, pcm); eng.drain()
st = O.Stream(om, 13); st.push(pcm)
assert np.array_equal(eng.pop_tokens(sid), st.tokens())
pcm6 = synth.synth_pcm(22, 6.0)                                                              # 76 frames on a 1-slot engine: the workspace grows
toks6, enc6, frames6 = eng.transcribe_full(pcm6, want_frames=True)
enc6_o, toks6_o, frames6_o = om.transcribe_full(O.Preproc(model=om).process(pcm6))
assert float(np.abs(enc6 - enc6_o).max() / np.abs(enc6_o).max()) < 1e-4
assert np.array_equal(toks6, toks6_o) and np.array_equal(frames6, frames6_o)
toks_again, _ = eng.transcribe_full(pcm)                                                      # and the shorter one again afterwards
assert np.array_equal(toks_again, toks_o)
eng.close()

e16 = nsb200.Engine(path, right_context=13, max_streams=1, compute=nsb200.COMPUTE_F16, kv_dtype=nsb200.KV_F16)
_, enc16 = e16.transcribe_full(pcm)
rel16 = float(np.abs(enc16 - enc_o).max() / np.abs(enc_o).max())
print("fp16: encoder rel err vs the f32 oracle", rel16)
assert rel16 < 1e-2, rel16
e16.close()
# the reference's own src/transcribe.cpp (byte-identical, oracle/Makefile) on the drop-in library: stdout carries the transcript
cli = os.path.join(ROOT, "oracle", "_ref", "nemotron-asr-batch-dropin")
if os.path.exists(cli):
    import subprocess, tempfile
    with tempfile.TemporaryDirectory() as d:
        f = os.path.join(d, "a.pcm"); pcm.tofile(f)
        r = subprocess.run([cli, path, f], capture_output=True, text=True, timeout=120, env=dict(os.environ, NSB_COMPUTE="1"))
    assert r.returncode == 0, r.stderr
    assert "=== Transcription ===\n" + om.detok(toks_o) + "\n" in r.stdout, r.stdout
    print("batch CLI drop-in ok")
# full depth, 30 s of audio (375 encoder frames, full-context attention over all of them)
path24 = synth.cached_model("f32", 24, R=13)
pcm30 = synth.synth_pcm(23, 30.0)
om24 = O.Model(path24)
enc_o, toks_o, frames_o = om24.transcribe_full(O.Preproc(model=om24).process(pcm30))
e24 = nsb200.Engine(path24, right_context=13, max_streams=1, compute=nsb200.COMPUTE_F32)
toks, enc, frames = e24.transcribe_full(pcm30, want_frames=True)
rel24 = float(np.abs(enc - enc_o).max() / np.abs(enc_o).max())
print("24 layers x 30 s, strict fp32: encoder rel err", rel24, "tokens", len(toks), len(toks_o))
assert enc.shape == enc_o.shape and rel24 < 3e-4, rel24
assert np.array_equal(toks, toks_o) and np.array_equal(frames, frames_o)
e24.close()
print("BATCH PATH OK")
"""


@pytest.mark.gpu
def test_batch_path_matches_checker_and_reference_fixture(built):
    r = subprocess.run([sys.executable, "-c", SCRIPT, ROOT], capture_output=True, text=True, timeout=900)
    sys.stdout.write(r.stdout[-2000:])
    assert r.returncode == 0 and "BATCH PATH OK" in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])
