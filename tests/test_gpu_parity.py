"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI (libnsb200.so), against the CPU oracle on the
same seeded inputs, against the committed golden vectors from the reference's own code, and through
size-independent properties. Tolerances (relative, max-norm over the tensor):
   f32 compute  : encoder / logits <= 1e-4, greedy tokens identical
   f16 compute  : encoder <= 3e-3 vs the oracle run with the same rounding points (north-star: ~1e-3), tokens identical
                  up to decisions whose top-2 logit gap is inside the numerical noise band
   bf16 compute : encoder <= 3e-2 (8-bit mantissa), same token rule
   Q8_0         : compared against the oracle's Q8_0 arithmetic (weights dequantised to fp16, as the fused kernel does)
"""
import os
import subprocess

import numpy as np
import pytest

import oracle as O
import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def rel(a, b):
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def run_engine_vs_oracle(eng, om, R, audio, read=None, taps=True):
    """Feed every stream in CLI-sized reads, step whenever something is ready; return per-stream engine tokens,
    oracle streams, and the worst encoder error over all (stream, chunk)."""
    import nsb200
    T = 1 + R
    n = len(audio)
    ids = [eng.open_stream() for _ in range(n)]
    orc = [O.Stream(om, R, trace=True) for _ in range(n)]
    for s in range(n):
        orc[s].push(audio[s])
    read = read or eng.chunk_samples
    pos, done, worst = [0] * n, [0] * n, 0.0
    while any(pos[s] < len(audio[s]) for s in range(n)):
        for s in range(n):
            if pos[s] < len(audio[s]):
                eng.push(ids[s], audio[s][pos[s]:pos[s] + read]); pos[s] += read
        while True:
            ready = [s for s in range(n) if eng.ready(ids[s])]
            if not ready:
                break
            assert eng.step() == len(ready)
            if taps:
                enc = eng.debug_get("enc", len(ready))
                for bi, s in enumerate(ready):
                    worst = max(worst, rel(enc[bi * T:(bi + 1) * T], orc[s].trace_enc(done[s])))
                    done[s] += 1
    toks = [eng.pop_tokens(ids[s]) for s in range(n)]
    for s in range(n):
        assert eng.chunks(ids[s]) == orc[s].chunks
    return toks, orc, worst, ids


def assert_tokens_match_up_to_near_ties(toks, orc, band, details=None):
    """Identical token sequences, except that a stream may diverge at a decision whose top-2 gap in the ORACLE is
    below `band` (numerical noise of 16-bit activations); nothing can be said after such a point (the LSTM state
    differs from there on). Returns the number of fully identical streams. `details` (a dict) receives what the comparison saw:
    decisions compared (joint evaluations of the oracle up to each stream's first divergence), the divergences with the smallest
    top-2 gap in their window, and how common gaps below `band` are among ALL oracle decisions (how much the rule lets through)."""
    identical = 0
    n_dec, n_small, n_all, flips = 0, 0, 0, []
    for s, (tg, o) in enumerate(zip(toks, orc)):
        to = o.tokens()
        if details is not None:
            for e in range(o.n_evals()):
                lg = np.sort(o.trace_logits(e)); n_all += 1; n_small += int(lg[-1] - lg[-2] < band)
        if len(tg) == len(to) and np.array_equal(tg, to):
            identical += 1
            n_dec += o.n_evals()
            continue
        # walk the oracle's evaluations to the first decision that differs
        k = next((i for i in range(min(len(tg), len(to))) if tg[i] != to[i]), min(len(tg), len(to)))
        emitted, ev = 0, 0
        while ev < o.n_evals():
            tok = o.eval_token(ev)
            if tok != 1024:
                if emitted == k:
                    break
                emitted += 1
            elif emitted == k and len(tg) > k:      # oracle said blank where the engine emitted
                pass
            ev += 1
        # the divergent decision is evaluation ev (or an earlier blank/non-blank flip right before it): accept if ANY of the
        # evaluations since the last agreed emission has a top-2 gap inside the band
        # ... or anywhere in the emission run that leads up to it: random joint weights emit runs of up to 10 symbols per frame,
        # often repeating a token, and one repetition more or fewer (a flip at a near-tie INSIDE the run) leaves the two
        # sequences identical until the run ends -- the first differing INDEX then sits later than the flipped DECISION.
        # So the window is the last two frames' worth of evaluations (2 x (10 symbols + 1 blank)).
        lo = ev
        while lo > 0 and o.eval_token(lo - 1) == 1024:
            lo -= 1
        lo = max(0, min(lo, ev - 22))
        gaps = []
        for e in range(lo, min(ev + 1, o.n_evals())):
            lg = np.sort(o.trace_logits(e)); gaps.append(float(lg[-1] - lg[-2]))
        n_dec += min(ev + 1, o.n_evals())
        flips.append({"stream": s, "token_index": k, "min_gap": min(gaps) if gaps else None})
        assert gaps and min(gaps) < band, f"stream {s}: tokens diverge at {k} with oracle top-2 gaps {gaps} (band {band})"
    if details is not None:
        details.update(decisions_compared=n_dec, divergences=flips, flip_rate=len(flips) / max(n_dec, 1),
                       fraction_of_oracle_decisions_with_gap_below_band=n_small / max(n_all, 1))
    return identical


# ------------------------------------------------------------------------------------------------------------
def test_logmel_kernel_parity(built):
    import nsb200
    path = synth.cached_model("f32", 2, R=0)
    eng = nsb200.Engine(path, right_context=0, max_streams=1, compute=nsb200.COMPUTE_F32)
    om = O.Model(path)
    pcm = np.stack([synth.synth_pcm(20 + i, 1.0) for i in range(5)])
    got = eng.op_logmel(pcm)
    for i in range(5):
        ref = O.Preproc(model=om).process(pcm[i])
        assert got[i].shape == ref.shape
        ulp = np.abs(got[i].view(np.int32).astype(np.int64) - ref.view(np.int32).astype(np.int64))
        assert ulp.max() <= 1 and np.mean(ulp == 0) > 0.99           # everything but the final log is bit-identical
    g = np.load(os.path.join(GOLD, "mel_ref.npz"))                   # the reference's own output
    ulp = np.abs(eng.op_logmel(g["pcm"])[0].view(np.int32).astype(np.int64) - g["mel"].view(np.int32).astype(np.int64))
    assert ulp.max() <= 1
    ulp = np.abs(eng.op_logmel(g["sine"])[0].view(np.int32).astype(np.int64) - g["mel_sine"].view(np.int32).astype(np.int64))
    assert ulp.max() <= 1
    # edge cases: silence, full-scale, shortest input that yields a frame, too-short input
    assert np.allclose(eng.op_logmel(np.zeros(256, np.int16))[0], np.log(np.float32(2.0 ** -24)))
    assert eng.op_logmel(np.zeros(255, np.int16)).shape[1] == 0
    assert np.isfinite(eng.op_logmel(np.full(3000, 32767, np.int16))).all()
    eng.close()


@pytest.mark.parametrize("wtype,compute,mm,fused", [("f32", 2, O.MM_F16, 0), ("f32", 3, O.MM_BF16, 0), ("q8_0", 0, O.MM_Q8FAST, 0), ("q4_0", 0, O.MM_Q8FAST, 0),
                                                     ("q8_0", 0, O.MM_Q8FAST, 1), ("q4_0", 0, O.MM_Q8FAST, 1)])
def test_tcgen05_gemm_parity(built, wtype, compute, mm, fused, monkeypatch):
    import nsb200
    if fused:                                      # the fused-dequantisation kernels (gemm_q8_kernel), otherwise behind the fp16 expansion
        monkeypatch.setenv("NSB_Q8_PREDEQUANT_ROWS", "100000")
    path = synth.cached_model(wtype, 2, R=0)
    eng = nsb200.Engine(path, right_context=0, max_streams=1, compute=compute)
    om = O.Model(path, mm)
    rng = np.random.default_rng(0)
    P = "encoder.layers.1."
    for name, k in (("feed_forward1.linear1.weight", 1024), ("feed_forward2.linear2.weight", 4096), ("conv.pointwise_conv1.weight", 1024)):
        for rows in (1, 7, 128, 129, 333):                            # M tails, multiple M tiles
            x = rng.standard_normal((rows, k)).astype(np.float32)
            assert rel(eng.op_gemm(P + name, x), om.matmul(P + name, x)) < 3e-5, (name, rows)
    x = rng.standard_normal((64, 1024)).astype(np.float32)           # linearity: W(ax + by) = aWx + bWy up to fp rounding
    y = rng.standard_normal((64, 1024)).astype(np.float32)
    n = P + "self_attn.linear_out.weight"
    lhs = eng.op_gemm(n, (x + y).astype(np.float32))
    assert rel(lhs, eng.op_gemm(n, x) + eng.op_gemm(n, y)) < 2e-2
    eng.close()


@pytest.mark.parametrize("persist", ["0", "1"])
@pytest.mark.parametrize("wtype,compute,mm", [("f32", 2, O.MM_F16), ("f32", 3, O.MM_BF16), ("q8_0", 0, O.MM_Q8FAST), ("q4_0", 0, O.MM_Q8FAST)])
def test_tcgen05_gemm_large_batch_pair_tiles(built, wtype, compute, mm, persist, monkeypatch):
    """>= 4 row tiles: 256-row CTA-pair tiles (cta_group::2, M = 256) with BN in {256, 208, 160, 112} chosen per shape -- every
    BN, the overhang of the last tile along N (4096 = 19 x 208 + 144), a ragged last row tile (900 = 3 x 256 + 132), and in
    Q8_0 mode the per-launch dequantisation into the fp16 scratch in front of it. persist = 1: the persistent variant of the same tiles
    (one CTA pair per SM pair walks the tiles, two TMEM accumulators, one TMA ring across tiles; 900 rows x 1024 columns = 4 x 10 tiles
    x 2 k-slices on 74 pairs: pairs with two tiles and pairs with one)."""
    import nsb200
    monkeypatch.setenv("NSB_PAIR256_PERSIST", persist)
    if wtype in ("q8_0", "q4_0"):
        # quantised weights: 0 = dequantisation fused into the CTA-pair tiles (gemm_q8_pair256_kernel, the default), 1 = the earlier path
        # (dequantise into an fp16 scratch per launch, then the fp16 pair tiles)
        monkeypatch.setenv("NSB_Q8_PAIR", "1" if persist == "0" else "0")
    path = synth.cached_model(wtype, 2, R=0)
    eng = nsb200.Engine(path, right_context=0, max_streams=1, compute=compute)
    om = O.Model(path, mm)
    rng = np.random.default_rng(1)
    P = "encoder.layers.1."
    for rows in (1792, 900):
        for name, k in (("feed_forward1.linear1.weight", 1024), ("feed_forward2.linear2.weight", 4096), ("conv.pointwise_conv1.weight", 1024),
                        ("self_attn.linear_qkv.weight", 1024)):
            x = rng.standard_normal((rows, k)).astype(np.float32)
            if "qkv" in name:
                ref = np.concatenate([om.matmul(P + f"self_attn.linear_{c}.weight", x) for c in "qkv"], axis=1)
            else:
                ref = om.matmul(P + name, x)
            y = eng.op_gemm(P + name, x)
            assert y.shape == ref.shape and rel(y, ref) < 3e-5, (name, rows)
    eng.close()


@pytest.mark.parametrize("R", [0, 1, 6, 13])
def test_streaming_parity_f32_all_latency_modes(built, R):
    import nsb200
    path = synth.cached_model("f32", 2, R=R)
    eng = nsb200.Engine(path, right_context=R, max_streams=3, compute=nsb200.COMPUTE_F32)
    eng.debug_enable(True)
    om = O.Model(path)
    audio = [synth.synth_pcm(s, 2.2 + 0.41 * s + 0.2 * R) for s in range(3)]        # ragged lengths: batch size varies per step
    toks, orc, worst, _ = run_engine_vs_oracle(eng, om, R, audio)
    assert worst < 1e-4, worst
    for s in range(3):
        assert np.array_equal(toks[s], orc[s].tokens()), s
    assert sum(len(t) for t in toks) > 0
    eng.close()


def test_long_form_stream_two_minutes_strict_fp32(built):
    """BASELINE.json config 5's long-form aspect at test size: two streams of 2 minutes at 1.12 s chunks (106 chunks each: the 70-row
    ring wraps 21 times, 1.9 M samples through the host-side staging with its lazy compaction), one fed in one-hour-style bulk (a single
    push, then drained), one in chunk-sized reads; strict fp32 with the CUDA graph on and no taps: tokens identical to the checker (reference: src/nemo-stream.cpp:961-1057, 1079-1127)."""
    import nsb200
    R, T = 13, 14
    path = synth.cached_model("f32", 2, R=R)
    eng = nsb200.Engine(path, right_context=R, max_streams=2, compute=nsb200.COMPUTE_F32, cuda_graph=True)
    om = O.Model(path)
    audio = [synth.synth_pcm(40 + s, 120.0) for s in range(2)]
    ids = [eng.open_stream() for _ in range(2)]
    eng.push(ids[0], audio[0])                                                   # bulk: everything at once
    pos, read = 0, eng.chunk_samples
    while pos < len(audio[1]) or eng.ready(ids[0]) or eng.ready(ids[1]):
        if pos < len(audio[1]):
            eng.push(ids[1], audio[1][pos:pos + read]); pos += read
        if eng.step() == 0 and pos >= len(audio[1]):
            break
    toks = [eng.pop_tokens(i) for i in ids]
    for s in range(2):
        o = O.Stream(om, R)
        o.push(audio[s])
        assert eng.chunks(ids[s]) == o.chunks >= 106, (eng.chunks(ids[s]), o.chunks)
        assert np.array_equal(toks[s], o.tokens()), s
        assert len(toks[s]) > 1000
    eng.close()


def test_streaming_parity_f32_full_24_layer_model(built):
    import nsb200
    path = synth.cached_model("f32", 24, R=1)
    eng = nsb200.Engine(path, right_context=1, max_streams=2, compute=nsb200.COMPUTE_F32)
    eng.debug_enable(True)
    om = O.Model(path)
    audio = [synth.synth_pcm(30 + s, 1.6 + 0.3 * s) for s in range(2)]
    toks, orc, worst, _ = run_engine_vs_oracle(eng, om, 1, audio)
    assert worst < 2e-4, worst
    for s in range(2):
        assert np.array_equal(toks[s], orc[s].tokens()), s
    eng.close()


@pytest.mark.parametrize("wtype,compute,kv,mm,okv,tol,band,R", [
    ("f32", 2, 0, O.MM_F16, O.KV_F32, 3e-3, 2e-2, 1),
    ("f16", 0, 1, O.MM_REF, O.KV_F16, 3e-3, 2e-2, 1),       # F16 GGUF -> ggml F16 semantics (+ fp16 K/V ring)
    ("f32", 3, 2, O.MM_BF16, O.KV_BF16, 3e-2, 2e-1, 1),
    ("q8_0", 0, 0, O.MM_Q8FAST, O.KV_F32, 3e-3, 2e-2, 1),
    ("q4_0", 0, 0, O.MM_Q8FAST, O.KV_F32, 3e-3, 2e-2, 1),    # Q4_0 GGUF: nibbles resident in HBM, dequantisation fused into the GEMM operand path
    ("q4_0", 2, 1, O.MM_Q8FAST, O.KV_F16, 3e-3, 2e-2, 1),    # the same file expanded to fp16 at load (NSB_COMPUTE_F16): same operand values
    ("f16", 0, 1, O.MM_REF, O.KV_F16, 3e-3, 2e-2, 0),       # 80 ms mode (T = 1): paired attention kernel, odd batches
    ("f16", 0, 1, O.MM_REF, O.KV_F16, 3e-3, 2e-2, 6),       # 560 ms mode (T = 7): tiled attention kernel with a 16-bit ring
])
def test_streaming_parity_16bit_and_q8(built, wtype, compute, kv, mm, okv, tol, band, R):
    import nsb200
    path = synth.cached_model(wtype, 2, R=R)
    eng = nsb200.Engine(path, right_context=R, max_streams=4, compute=compute, kv_dtype=kv)
    eng.debug_enable(True)
    om = O.Model(path, mm, okv)
    audio = [synth.synth_pcm(50 + s, 2.0 + 0.3 * s) for s in range(4)]
    toks, orc, worst, _ = run_engine_vs_oracle(eng, om, R, audio)
    assert worst < tol, worst
    identical = assert_tokens_match_up_to_near_ties(toks, orc, band)
    assert identical >= 2, identical
    eng.close()


@pytest.mark.parametrize("wtype", ["q8_0", "q4_0"])
def test_streaming_parity_quantised_fused_operand_path(built, wtype, monkeypatch):
    """The fused-dequantisation GEMM kernels inside the step (gemm_q8_kernel: raw quants by TMA, dequantiser warps, tcgen05) -- since the
    layer-ahead fp16 shadows became the default at every batch size they run only on request (NSB_Q8_PREDEQUANT_ROWS above the batch's
    rows): same operand values, so the same tolerances as the default path."""
    import nsb200
    monkeypatch.setenv("NSB_Q8_PREDEQUANT_ROWS", "100000")
    R = 1
    path = synth.cached_model(wtype, 2, R=R)
    eng = nsb200.Engine(path, right_context=R, max_streams=4, compute=0, kv_dtype=0)
    eng.debug_enable(True)
    om = O.Model(path, O.MM_Q8FAST, O.KV_F32)
    audio = [synth.synth_pcm(20 + s, 2.0 + 0.3 * s) for s in range(4)]
    toks, orc, worst, _ = run_engine_vs_oracle(eng, om, R, audio)
    assert worst < 3e-3, worst
    assert assert_tokens_match_up_to_near_ties(toks, orc, 2e-2) >= 2
    eng.close()


@pytest.mark.parametrize("wtype,compute,kv,mm,okv,tol,band", [
    ("f16", 0, 1, O.MM_REF, O.KV_F16, 3e-3, 2e-2),
    ("f32", 3, 2, O.MM_BF16, O.KV_BF16, 3e-2, 2e-1),
    ("q8_0", 0, 0, O.MM_Q8FAST, O.KV_F32, 3e-3, 2e-2),
])
def test_large_batch_in_engine_parity(built, wtype, compute, kv, mm, okv, tol, band):
    """128 streams x 1.12 s chunks = 1792 token rows per step: the large-batch kernels INSIDE the step (256-row CTA-pair GEMM
    tiles with SiLU / 16-bit, fp32 and residual epilogues, Q8_0 pre-dequantisation, 3-CTA-per-SM attention at T = 14) against
    the oracle on three of the streams, and batch invariance over the rest (streams fed the same audio give the same tokens)."""
    import nsb200
    R, n = 13, 128
    T = 1 + R
    path = synth.cached_model(wtype, 2, R=R)
    eng = nsb200.Engine(path, right_context=R, max_streams=n, compute=compute, kv_dtype=kv)
    eng.debug_enable(True)
    om = O.Model(path, mm, okv)
    base = [synth.synth_pcm(300 + s, 2.7) for s in range(8)]
    audio = np.stack([base[s % 8] for s in range(n)])
    ids = [eng.open_stream() for _ in range(n)]
    eng.push_batch(ids, audio)
    orc = [O.Stream(om, R, trace=True) for _ in range(3)]
    for s in range(3):
        orc[s].push(audio[s])
    worst, chunk = 0.0, 0
    while eng.ready(ids[0]):
        assert eng.step() == n
        enc = eng.debug_get("enc", n)
        for s in range(3):
            worst = max(worst, rel(enc[s * T:(s + 1) * T], orc[s].trace_enc(chunk)))
        for s in range(8, n):                                            # same audio, another batch row: same encoder rows
            if s % 41 == 0:
                assert np.array_equal(enc[s * T:(s + 1) * T], enc[(s % 8) * T:(s % 8 + 1) * T]), s
        chunk += 1
    assert chunk == orc[0].chunks >= 2
    assert worst < tol, worst
    toks = [eng.pop_tokens(i) for i in ids]
    assert_tokens_match_up_to_near_ties(toks[:3], orc, band)
    for s in range(8, n):
        assert np.array_equal(toks[s], toks[s % 8]), s
    eng.close()


def test_q8_fast_vs_ggml_q8_semantics_delta_is_small(built):
    """The fused-dequant kernel keeps activations in fp16 (more accurate than ggml, which quantises activations to Q8_0
    too): report/limit the distance to the reference's Q8_0 arithmetic."""
    import nsb200
    path = synth.cached_model("q8_0", 2, R=1)
    eng = nsb200.Engine(path, right_context=1, max_streams=1, compute=0)
    eng.debug_enable(True)
    om = O.Model(path, O.MM_REF)                                        # activations quantised per 32, integer block dots
    audio = [synth.synth_pcm(77, 1.5)]
    _, _, worst, _ = run_engine_vs_oracle(eng, om, 1, audio)
    assert worst < 3e-2, worst
    eng.close()


def test_first_chunk_matches_reference_golden(built):
    """tests/golden/model_ref_L2.npz was produced by the reference's src/reference code: first chunk of an R=1 stream."""
    import nsb200
    g = np.load(os.path.join(GOLD, "model_ref_L2.npz"))
    pcm = np.load(os.path.join(GOLD, "mel_ref.npz"))["pcm"]
    eng = nsb200.Engine(synth.cached_model("f32", 2, R=0), right_context=1, max_streams=1, compute=nsb200.COMPUTE_F32)
    eng.debug_enable(True)
    s = eng.open_stream()
    eng.push(s, pcm[:160 * 15 + 256])                                   # exactly one chunk's worth
    assert eng.step() == 1 and eng.step() == 0
    assert rel(eng.debug_get("mel", 1)[0], g["chunk"]) < 1e-6
    assert rel(eng.debug_get("sub", 1), g["sub"][2:]) < 1e-4
    assert rel(eng.debug_get("layer.0", 1), g["layer0"]) < 1e-4
    assert rel(eng.debug_get("enc", 1), g["layer1"]) < 1e-4
    assert rel(eng.debug_get("logits", 1)[0], g["logits0"]) < 1e-4
    assert np.array_equal(eng.pop_tokens(s), g["tokens"])
    eng.close()


@pytest.mark.parametrize("R,secs", [(0, 6.0), (1, 6.6), (6, 7.5), (13, 9.0)])
def test_cached_streaming_matches_reference_golden(built, R, secs):
    """tests/golden/cached_ref_L2.npz: whole streams run past the roll of the 70-row cache on the reference's own compiled
    modules (tools/make_golden.py section 3). The CUDA path, strict fp32: identical greedy tokens, last chunk's encoder output
    <= 1e-4 relative."""
    import nsb200
    g = np.load(os.path.join(GOLD, "cached_ref_L2.npz"))
    eng = nsb200.Engine(synth.cached_model("f32", 2, R=0), right_context=R, max_streams=1, compute=nsb200.COMPUTE_F32)
    eng.debug_enable(True)
    s = eng.open_stream()
    eng.push(s, synth.synth_pcm(11, secs))
    assert eng.drain() == int(g[f"chunks_R{R}"]) == eng.chunks(s)
    assert rel(eng.debug_get("enc", 1), g[f"enc_last_R{R}"]) < 1e-4
    assert np.array_equal(eng.pop_tokens(s), g[f"tokens_R{R}"])
    eng.close()


def test_ring_cache_equals_rolled_cache_and_reset(built):
    import nsb200
    R = 6
    path = synth.cached_model("f32", 2, R=R)
    eng = nsb200.Engine(path, right_context=R, max_streams=2, compute=nsb200.COMPUTE_F32)
    om = O.Model(path)
    audio = [synth.synth_pcm(60, 7.3), synth.synth_pcm(61, 3.1)]
    toks, orc, _, ids = run_engine_vs_oracle(eng, om, R, audio, taps=False)
    for s in range(2):
        for which in (0, 1, 2):                                           # K ring (logical order), V ring, conv state
            for layer in (0, 1):
                assert rel(eng.debug_cache(ids[s], which, layer), orc[s].cache(which, layer)) < 1e-4, (s, which, layer)
    # reset re-zeroes the device caches (the reference's reset leaves them stale): same audio -> same tokens
    eng.reset_stream(ids[1])
    eng.push(ids[1], audio[1]); eng.drain()
    assert np.array_equal(eng.pop_tokens(ids[1]), toks[1])
    # slot reuse after close
    eng.close_stream(ids[0])
    s2 = eng.open_stream()
    assert s2 == ids[0]
    eng.push(s2, audio[1]); eng.drain()
    assert np.array_equal(eng.pop_tokens(s2), toks[1])
    eng.close()


def test_api_edge_cases(built):
    import nsb200
    path = synth.cached_model("f32", 2, R=0)
    eng = nsb200.Engine(path, right_context=0, max_streams=2, compute=nsb200.COMPUTE_F32)
    om = O.Model(path)
    assert eng.step() == 0                                                # nothing open / ready
    s = eng.open_stream()
    eng.push(s, np.zeros(0, np.int16))                                    # empty push is a no-op (reference returns "")
    assert not eng.ready(s) and eng.step() == 0 and len(eng.pop_tokens(s)) == 0
    pcm = synth.synth_pcm(70, 1.0)
    for i in range(0, 1500):                                              # one sample at a time across the first chunk boundary
        eng.push(s, pcm[i:i + 1])
    eng.push(s, pcm[1500:])
    n = eng.drain()
    ref = O.Stream(om, 0); ref.push(pcm)
    assert n == ref.chunks and np.array_equal(eng.pop_tokens(s), ref.tokens())
    with pytest.raises(nsb200.NsbError):
        eng.push(1, pcm)                                                  # slot 1 not open
    with pytest.raises(nsb200.NsbError):
        eng.push(99, pcm)
    eng.open_stream()
    with pytest.raises(nsb200.NsbError, match="no free stream slot"):
        eng.open_stream()
    assert eng.detok(ref.tokens()) == om.detok(ref.tokens())
    eng.close()


@pytest.mark.parametrize("seed,R", [(11, 1), (12, 0), (13, 6)])
def test_randomised_serving_schedule_matches_per_session_oracle(built, seed, R):
    """A randomised serving schedule over 5 stream slots: pushes of random size (1 sample .. 2.5 chunks, now and then nothing), plain
    steps and split steps with up to three in flight, streams that end (close -> the slot is reused by a new session), streams reset
    in mid-utterance, slots that stay idle for a while -- so the batch composition, the batch size and every slot's cache fill level
    change from step to step. Strict fp32: every SESSION's tokens must equal the checker's run over exactly the audio that session
    pushed since its open / reset, and the chunk counts must agree (reference: the per-stream driver, src/nemo-stream.cpp:1079-1127,
    one nemo_stream_context per session)."""
    import nsb200
    rng = np.random.default_rng(seed)
    T = 1 + R
    path = synth.cached_model("f32", 2, R=R)
    eng = nsb200.Engine(path, right_context=R, max_streams=5, compute=nsb200.COMPUTE_F32, cuda_graph=True)
    om = O.Model(path)
    shift = eng.shift_samples
    pool = [synth.synth_pcm(400 + 10 * seed + k, 4.0 + 0.7 * k) for k in range(6)]
    sessions = {}                                            # slot -> dict(audio=recording, pos, pushed=list of arrays, toks=list)
    finished = []                                            # (pushed audio, tokens, chunks)
    inflight = 0

    def open_session():
        nonlocal inflight
        slot = eng.open_stream()                             # collects the steps in flight first (nsb200.h)
        inflight = 0
        sessions[slot] = dict(audio=pool[int(rng.integers(len(pool)))], pos=0, pushed=[], toks=[])

    def harvest(slot):
        sessions[slot]["toks"] += eng.pop_tokens(slot).tolist()

    def finish(slot, close):
        nonlocal inflight
        ses = sessions[slot]
        # close / reset collect the steps in flight first (nsb200.h); ready chunks that were never launched are dropped with the session,
        # so run them now: the checker sees everything that was pushed
        while inflight:
            eng.step_end(); inflight -= 1
        while eng.ready(slot):
            eng.step()
        harvest(slot)
        finished.append((np.concatenate(ses["pushed"]) if ses["pushed"] else np.zeros(0, np.int16), ses["toks"], eng.chunks(slot)))
        if close:
            eng.close_stream(slot); del sessions[slot]
        else:
            eng.reset_stream(slot); sessions[slot] = dict(audio=pool[int(rng.integers(len(pool)))], pos=0, pushed=[], toks=[])

    for _ in range(3):
        open_session()
    for tick in range(140):
        for slot in list(sessions):
            ses = sessions[slot]
            u = rng.random()
            if ses["pos"] >= len(ses["audio"]):
                finish(slot, close=True)
                continue
            if u < 0.03:
                finish(slot, close=False)                    # reset in mid-utterance
                continue
            if u < 0.15:
                continue                                     # this stream is idle for a tick
            n = int(rng.choice([1, 17, shift // 3, shift, int(2.5 * shift)]))
            piece = ses["audio"][ses["pos"]:ses["pos"] + n]
            eng.push(slot, piece); ses["pushed"].append(piece); ses["pos"] += len(piece)
        if len(sessions) < 5 and rng.random() < 0.2:
            open_session()
        mode = rng.random()
        if mode < 0.4:
            while inflight:
                assert eng.step_end() > 0; inflight -= 1
            eng.step()
        else:
            while inflight < 3 and eng.step_begin() > 0:
                inflight += 1
            if inflight and rng.random() < 0.7:
                assert eng.step_end() > 0; inflight -= 1
        for slot in sessions:
            harvest(slot)
    for slot in list(sessions):
        finish(slot, close=True)
    assert len(finished) >= 6
    n_tok = 0
    for pushed, toks, chunks in finished:
        o = O.Stream(om, R)
        if len(pushed):
            o.push(pushed)
        assert chunks == o.chunks, (chunks, o.chunks)
        assert toks == o.tokens().tolist()
        n_tok += len(toks)
    assert n_tok > 50
    eng.close()


def test_loader_rejects_mis_shaped_and_mis_typed_matrices_by_name(built, tmp_path):
    """Tensor-type / shape validation of the per-layer matrices (the reference's loader checks presence only, nemo-ggml.cpp:362-384):
    a [4096, 1024] matrix stored as [1024, 4096], and a matrix of an unsupported ggml type, are refused with the tensor's name."""
    import struct
    import nsb200
    raw = bytearray(open(synth.cached_model("f16", 2, R=0), "rb").read())
    name = b"encoder.layers.1.feed_forward1.linear2.weight"
    at = raw.index(struct.pack("<Q", len(name)) + name) + 8 + len(name)
    nd, d0, d1, typ = struct.unpack_from("<IqqI", raw, at)
    assert (nd, d0, d1, typ) == (2, 4096, 1024, 1)
    bad = bytearray(raw); struct.pack_into("<Iqq", bad, at, 2, 1024, 4096)
    f = tmp_path / "swapped.gguf"; f.write_bytes(bad)
    with pytest.raises(nsb200.NsbError, match=r"feed_forward1.linear2.weight' has shape \[1024, 4096\], expected \[4096, 1024\]"):
        nsb200.Engine(str(f), right_context=0, max_streams=1)
    bad = bytearray(raw); struct.pack_into("<I", bad, at + 20, 12)         # type 12 = Q4_K: not a type the converter writes
    f = tmp_path / "q4k.gguf"; f.write_bytes(bad)
    with pytest.raises(nsb200.NsbError, match="feed_forward1.linear2.weight' has unsupported type 12"):
        nsb200.Engine(str(f), right_context=0, max_streams=1)


def test_many_streams_batch_invariance(built):
    """A stream's tokens must not depend on which other streams share its batch (64 streams vs alone)."""
    import nsb200
    R = 1
    path = synth.cached_model("f32", 2, R=R)
    eng = nsb200.Engine(path, right_context=R, max_streams=64, compute=nsb200.COMPUTE_F32)
    audio = [synth.synth_pcm(100 + (s % 8), 1.2 + 0.05 * (s % 5)) for s in range(64)]
    ids = [eng.open_stream() for _ in range(64)]
    for s in range(64):
        eng.push(ids[s], audio[s])
    eng.drain()
    toks = [eng.pop_tokens(i) for i in ids]
    solo = nsb200.Engine(path, right_context=R, max_streams=1, compute=nsb200.COMPUTE_F32)
    for s in (0, 13, 63):
        sid = solo.open_stream(); solo.push(sid, audio[s]); solo.drain()
        assert np.array_equal(solo.pop_tokens(sid), toks[s]), s
        solo.close_stream(sid)
    for s in range(8, 64):
        if len(audio[s]) == len(audio[s - 8]):
            assert np.array_equal(toks[s], toks[s - 8])
    eng.close(); solo.close()


def test_maximum_batch_1024_streams_in_one_engine(built):
    """The largest batch one engine takes (the decode kernel's stream table holds 1024): 1024 streams x 80 ms chunks per step, bf16,
    graph on -- BASELINE.json config 4's whole population on one GPU. The streams carry 8 distinct recordings: rows with the same
    recording must produce the same tokens bit for bit (batch invariance up to the largest grid sizes), the first 8 rows must match the
    checker up to its near-ties, and a 1025th stream is refused."""
    import nsb200
    R, n = 0, 1024
    path = synth.cached_model("f16", 2, R=R)
    base = [synth.synth_pcm(300 + s, 1.0) for s in range(8)]
    L = min(len(b) for b in base)
    eng = nsb200.Engine(path, right_context=R, max_streams=n, compute=nsb200.COMPUTE_BF16, kv_dtype=nsb200.KV_BF16, cuda_graph=True)
    ids = np.array([eng.open_stream() for _ in range(n)], dtype=np.int32)
    with pytest.raises(nsb200.NsbError):
        eng.open_stream()
    eng.push_batch(ids, np.stack([base[s % 8][:L] for s in range(n)]))
    assert eng.drain() > 0
    toks = [eng.pop_tokens(int(i)) for i in ids]
    assert sum(len(t) for t in toks[:8]) > 20
    for s in range(8, n):
        assert np.array_equal(toks[s], toks[s % 8]), s
    om = O.Model(path, O.MM_BF16, O.KV_BF16)
    orc = []
    for s in range(8):
        o = O.Stream(om, R, trace=True); o.push(base[s][:L]); orc.append(o)
        assert eng.chunks(int(ids[s])) == o.chunks
    assert_tokens_match_up_to_near_ties(toks[:8], orc, 2e-1)
    eng.close()


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "nemotron-asr-dropin")), reason="drop-in CLI not prebuilt")
def test_reference_cli_dropin_end_to_end(built, tmp_path):
    """The reference's own src/transcribe_stream.cpp (byte-identical, compiled against include/) driving the engine:
    stdout = incremental pieces + full transcript + newline (transcribe_stream.cpp:156-174)."""
    path = synth.cached_model("f32", 2, R=13)
    pcm = synth.synth_pcm(5, 4.0)
    f = tmp_path / "a.pcm"
    pcm.tofile(f)
    exe = os.path.join(ROOT, "oracle", "_ref", "nemotron-asr-dropin")
    r = subprocess.run([exe, path, str(f), "70", "13"], capture_output=True, timeout=120)
    assert r.returncode == 0, r.stderr.decode()
    om = O.Model(path)
    ref = O.Stream(om, 13); ref.push(pcm)
    text = om.detok(ref.tokens())
    assert r.stdout.decode() == text + text + "\n"
    assert f"Chunks processed:    {ref.chunks}" in r.stderr.decode()


def test_batch_push_pop_equals_per_stream_calls(built):
    """nsb_push_pcm_batch / nsb_pop_tokens_batch (one C call per serving tick) == the per-stream entry points."""
    import nsb200
    R = 1
    path = synth.cached_model("f32", 2, R=R)
    audio = np.stack([synth.synth_pcm(120 + s, 1.5) for s in range(6)])
    a = nsb200.Engine(path, right_context=R, max_streams=6, compute=nsb200.COMPUTE_F32)
    ids = [a.open_stream() for _ in range(6)]
    for s in range(6):
        a.push(ids[s], audio[s])
    a.drain()
    ref = [a.pop_tokens(i) for i in ids]
    b = nsb200.Engine(path, right_context=R, max_streams=6, compute=nsb200.COMPUTE_F32)
    ids2 = [b.open_stream() for _ in range(6)]
    got = [[] for _ in range(6)]
    step = b.shift_samples
    for p in range(0, audio.shape[1], step):
        b.push_batch(ids2, audio[:, p:p + step])
        while b.step() > 0:
            pass
        toks, cnt = b.pop_tokens_batch(ids2, 64)
        for s in range(6):
            got[s] += toks[s, :cnt[s]].tolist()
    for s in range(6):
        assert got[s] == ref[s].tolist(), s
    with pytest.raises(nsb200.NsbError):
        b.push_batch([0, 99], audio[:2])                                 # bad stream id in the batch
    a.close(); b.close()


@pytest.mark.gpu
def test_split_step_overlapping_the_next_push_equals_plain_step(built):
    """nsb_engine_step_begin / _end with the next chunk pushed while the step is in flight == nsb_engine_step."""
    import nsb200
    R = 1
    path = synth.cached_model("f32", 2, R=R)
    audio = np.stack([synth.synth_pcm(140 + s, 1.3) for s in range(5)])
    a = nsb200.Engine(path, right_context=R, max_streams=5, compute=nsb200.COMPUTE_F32)
    ids = [a.open_stream() for _ in range(5)]
    for s in range(5):
        a.push(ids[s], audio[s])
    a.drain()
    ref = [a.pop_tokens(i) for i in ids]
    b = nsb200.Engine(path, right_context=R, max_streams=5, compute=nsb200.COMPUTE_F32)
    ids2 = [b.open_stream() for _ in range(5)]
    step = b.shift_samples
    assert b.step_end() == 0                                             # nothing in flight
    b.push_batch(ids2, audio[:, :step])
    pos = step
    got = [[] for _ in range(5)]
    while True:
        n = b.step_begin()
        if pos < audio.shape[1]:
            b.push_batch(ids2, audio[:, pos:pos + step]); pos += step    # overlaps the step in flight
        if n:
            assert b.step_end() == n
        toks, cnt = b.pop_tokens_batch(ids2, 64)
        for s in range(5):
            got[s] += toks[s, :cnt[s]].tolist()
        if n == 0 and pos >= audio.shape[1]:
            break
    for s in range(5):
        assert got[s] == ref[s].tolist(), s
        assert b.chunks(ids2[s]) == a.chunks(ids[s])
    a.close(); b.close()


@pytest.mark.gpu
@pytest.mark.parametrize("depth", [2, 3])
def test_steps_in_flight_equal_plain_steps(built, depth):
    """begin, begin, [begin,] end, begin, end, ...: steps i+1 (and i+2) are staged and enqueued while step i runs (the host side of a
    step and the device-side hand-over buffers are rings of three); tokens, their order and the chunk counts equal the
    one-step-at-a-time run; a fourth begin is an error; ready() counts launched chunks, chunks() collected ones; reset while steps
    are in flight collects them first."""
    import nsb200
    R = 1
    path = synth.cached_model("f32", 2, R=R)
    audio = np.stack([synth.synth_pcm(150 + s, 1.5) for s in range(4)])
    a = nsb200.Engine(path, right_context=R, max_streams=4, compute=nsb200.COMPUTE_F32)
    ids = [a.open_stream() for _ in range(4)]
    for s in range(4):
        a.push(ids[s], audio[s])
    a.drain()
    ref = [a.pop_tokens(i) for i in ids]
    b = nsb200.Engine(path, right_context=R, max_streams=4, compute=nsb200.COMPUTE_F32)
    ids2 = [b.open_stream() for _ in range(4)]
    b.push_batch(ids2, audio)                                            # everything up front: many chunks ready
    got = [[] for _ in range(4)]
    def collect():
        toks, cnt = b.pop_tokens_batch(ids2, 64)
        for s in range(4):
            got[s] += toks[s, :cnt[s]].tolist()
    n_launched = inflight = 0
    exhausted = False
    while True:
        while not exhausted and inflight < depth:
            n = b.step_begin()
            if n == 0:
                exhausted = True
                break
            assert n == 4
            n_launched += 1; inflight += 1
            if n_launched == 1:
                assert b.chunks(ids2[0]) == 0                            # launched, not collected
        if inflight == 3 and not exhausted:
            with pytest.raises(nsb200.NsbError):
                b.step_begin()                                           # at most three
        if inflight == 0:
            break
        assert b.step_end() == 4                                         # the OLDEST one
        inflight -= 1
        collect()
    assert b.step_end() == 0
    for s in range(4):
        assert got[s] == ref[s].tolist(), s
        assert b.chunks(ids2[s]) == a.chunks(ids[s]) == n_launched
    # reset with steps in flight: collected first, then the slot starts over
    b.reset_stream(ids2[0]); b.push(ids2[0], audio[0])
    for _ in range(depth):
        assert b.step_begin() == 1
    b.reset_stream(ids2[0])
    assert b.chunks(ids2[0]) == 0 and b.step_end() == 0
    b.push(ids2[0], audio[0]); b.drain()
    assert b.pop_tokens(ids2[0]).tolist() == ref[0].tolist()
    a.close(); b.close()


@pytest.mark.gpu
@pytest.mark.parametrize("wtype,compute,kv,mm,okv,bound", [
    ("f16", 0, 1, O.MM_REF, O.KV_F16, 0.03),
    ("f32", 3, 2, O.MM_BF16, O.KV_BF16, 0.3),
])
def test_decision_level_parity_16bit(built, wtype, compute, kv, mm, okv, bound):
    """Decision by decision (the logits tap of batch row 0): until the first flipped decision every joint evaluation of the
    engine is within `bound` of the oracle's logits, and a decision may only flip where the oracle's own top-2 gap is inside
    twice the distance measured so far -- the check the token-sequence comparison above cannot make exactly."""
    import nsb200
    R, T = 1, 2
    path = synth.cached_model(wtype, 2, R=R)
    eng = nsb200.Engine(path, right_context=R, max_streams=1, compute=compute, kv_dtype=kv)
    eng.debug_enable(True)
    om = O.Model(path, mm, okv)
    pcm = synth.synth_pcm(51, 2.3)
    o = O.Stream(om, R, trace=True); o.push(pcm)
    sid = eng.open_stream()
    pos, ev0, worst, n_cmp, flipped = 0, 0, 0.0, 0, False
    while pos < len(pcm) and not flipped:
        eng.push(sid, pcm[pos:pos + eng.chunk_samples]); pos += eng.chunk_samples
        while eng.ready(sid) and not flipped:
            assert eng.step() == 1
            lg = eng.debug_get("logits", 1)
            for i in range(lg.shape[0]):
                if ev0 + i >= o.n_evals():
                    break
                ol = o.trace_logits(ev0 + i)
                d = float(np.abs(lg[i] - ol).max()); worst = max(worst, d); n_cmp += 1
                assert d < bound, (ev0 + i, d)
                if int(np.argmax(lg[i])) != int(np.argmax(ol)):
                    srt = np.sort(ol)
                    assert srt[-1] - srt[-2] < 2 * worst + 1e-6, (ev0 + i, float(srt[-1] - srt[-2]), worst)
                    flipped = True
                    break
            ev0 += lg.shape[0]
    assert n_cmp >= 20, n_cmp
    eng.close()
