"""Non-streaming batch path (nsb_transcribe_full) on the GPU against its checker.

EXPERIMENTAL: the CUDA side (Engine::transcribe_full, attention_full_kernel, the un-chunked stem variant) was written after this
round's GPU budget was spent, so it has not run on hardware yet. The case therefore runs in a SUBPROCESS (a faulting kernel cannot
take the CUDA context of the parity suite with it), last in the suite, and is marked xfail(strict=False): XPASS = the path is
green against the oracle and the fixture produced by the reference's compiled modules; XFAIL = round-2 work, nothing else is
affected. The checker side is pinned on CPU (tests/test_oracle.py::test_batch_path_matches_compiled_reference)."""
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
enc_o, toks_o, _ = om.transcribe_full(mel)
g = np.load(os.path.join(ROOT, "tests", "golden", "batch_ref_L2.npz"))

eng = nsb200.Engine(path, right_context=13, max_streams=4, compute=nsb200.COMPUTE_F32)      # 4 x 14 = 56 workspace rows >= 39 frames
toks, enc = eng.transcribe_full(pcm)
assert enc.shape == enc_o.shape, (enc.shape, enc_o.shape)
rel = float(np.abs(enc - enc_o).max() / np.abs(enc_o).max())
print("strict fp32: encoder rel err vs oracle", rel, "tokens", len(toks), len(toks_o))
assert rel < 1e-4, rel
assert float(np.abs(enc[::4] - g["enc_every4"]).max() / np.abs(g["enc_every4"]).max()) < 1e-4
assert np.array_equal(toks, toks_o) and np.array_equal(toks, g["tokens"])
# the slot and the step workspace were only borrowed: a streaming stream on the same engine still matches the oracle
sid = eng.open_stream(); eng.push(sid, pcm); eng.drain()
st = O.Stream(om, 13); st.push(pcm)
assert np.array_equal(eng.pop_tokens(sid), st.tokens())
try:
    eng.transcribe_full(synth.synth_pcm(22, 6.0))                                             # 76 frames > 56 rows
    raise SystemExit("expected the workspace check to refuse")
except nsb200.NsbError as e:
    assert "do not fit" in str(e), e
eng.close()

e16 = nsb200.Engine(path, right_context=13, max_streams=4, compute=nsb200.COMPUTE_F16, kv_dtype=nsb200.KV_F16)
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
print("BATCH PATH OK")
"""


@pytest.mark.gpu
def test_batch_path_matches_checker_and_reference_fixture(built):
    r = subprocess.run([sys.executable, "-c", SCRIPT, ROOT], capture_output=True, text=True, timeout=240)
    sys.stdout.write(r.stdout[-2000:])
    assert r.returncode == 0 and "BATCH PATH OK" in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])
