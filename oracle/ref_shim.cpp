// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE. A thin extern "C" face over the reference's OWN
// code, compiled where it lies under /root/reference (never copied): src/preprocessor.cpp (the
// product log-mel front-end) and src/reference/*.cpp (the naive f32 model the reference's own
// tests use as their oracle). Built by oracle/Makefile into oracle/_ref/libnemo_ref.so and used
// by tests to pin oracle/stream_oracle.cpp and to generate tests/golden/*.
//
// Nothing here restates arithmetic; every call forwards to a reference class/function (ref_cached_layer_step only chooses the
// rows the reference's modules are run on).
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "preprocessor.h"          // /root/reference/src/preprocessor.h
#include "conformer_encoder.h"     // /root/reference/src/reference/include/*
#include "greedy_decode.h"
#include "rnnt_decoder.h"
#include "rnnt_joint.h"

using namespace nemo;

extern "C" {

// ---- src/preprocessor.cpp ----
void* ref_pp_new(const float* fb, const float* win400) {
    return nemo_preprocessor_init_from_data(fb, 128 * 257, win400, 400);
}
void ref_pp_free(void* p) { nemo_preprocessor_free((nemo_preprocessor*)p); }
int ref_pp_process(void* p, const int16_t* pcm, int n, float* out, int cap_frames) {
    std::vector<float> mel;
    size_t nf = nemo_preprocessor_process((nemo_preprocessor*)p, pcm, (size_t)n, mel);
    if ((int)nf > cap_frames) return -(int)nf;
    if (nf) memcpy(out, mel.data(), mel.size() * sizeof(float));
    return (int)nf;
}

// ---- src/reference: weights ("NEMO" v1 bin) ----
void* ref_weights_load(const char* path) {
    auto* w = new ModelWeights();
    if (!w->load(path)) { delete w; return nullptr; }
    return w;
}
void ref_weights_free(void* w) { delete (ModelWeights*)w; }

// ConvSubsampling::forward on one mel chunk [M,128] -> [t3,1024]
int ref_subsampling(void* w, const float* mel, int M, float* out, int cap_rows) {
    ConvSubsampling sub; sub.load_weights(*(ModelWeights*)w);
    TensorF in({1, (size_t)M, 128}); memcpy(in.ptr(), mel, (size_t)M * 128 * 4);
    TensorF o; sub.forward(in, o);
    int rows = (int)o.shape[1]; if (rows > cap_rows) return -rows;
    memcpy(out, o.ptr(), o.numel() * 4); return rows;
}

// ConformerLayer::forward (non-cached) on x [T,1024] with the reference's own pos table for T
int ref_layer_forward(void* w, int layer, const float* x, int T, float* y) {
    ConformerLayer L; L.load_weights(*(ModelWeights*)w, "encoder.layers." + std::to_string(layer));
    RelPositionalEncoding pe; TensorF pos; pe.get_pos_emb((size_t)T, pos);
    TensorF in({1, (size_t)T, 1024}); memcpy(in.ptr(), x, (size_t)T * 1024 * 4);
    TensorF o; L.forward(in, pos, o);
    memcpy(y, o.ptr(), o.numel() * 4); return T;
}

// One CACHED streaming step of a conformer layer, assembled only from the reference's own compiled modules (layer_norm,
// ConformerFeedForward, RelPositionMultiHeadAttention, ConformerConvolution -- the calls ConformerLayer::forward makes,
// conformer_encoder.cpp:29-69) so that the cached arithmetic of nemo-stream.cpp:435-545 / :308-384 can be pinned without ggml.
// The only thing added here is WHICH rows the modules see -- the caching itself, restated as windows:
//   attention: K / V are row-wise functions of LN_attn(x1), the cache keeps the last 70 rows (nemo-stream.cpp:477-484) and the
//              positional term depends on query - key distance only, so the T queries of this chunk over [cache | chunk] = the
//              last T rows of the non-cached module run on [att_hist (H <= 70 rows) | this chunk] with its own 2(H+T)-1 table;
//   conv:      the cache keeps the last 8 GLU rows (:368-381), a row-wise function of LN_conv(x2); the module pads 8 zero rows
//              on the left (= the zero-initialised cache), so conv over [conv_hist (Hc <= 8 rows) | chunk], last T rows.
// att_new / conv_new return this chunk's LN_attn(x1) / LN_conv(x2) rows for the caller's histories.
int ref_cached_layer_step(void* w, int layer, const float* x, int T, const float* att_hist, int H, const float* conv_hist, int Hc,
                          float* y, float* att_new, float* conv_new) {
    ConformerLayer L; L.load_weights(*(ModelWeights*)w, "encoder.layers." + std::to_string(layer));
    const size_t D = 1024, t = (size_t)T, n = t * D;
    TensorF r({1, t, D}), a, b;
    memcpy(r.ptr(), x, n * 4);
    layer_norm(r, L.norm_ff1_weight, L.norm_ff1_bias, D, 1e-5f, a);
    L.ffn1.forward(a, b);
    for (size_t i = 0; i < n; i++) r.data[i] += 0.5f * b.data[i];

    layer_norm(r, L.norm_attn_weight, L.norm_attn_bias, D, 1e-5f, a);
    memcpy(att_new, a.ptr(), n * 4);
    TensorF win({1, (size_t)H + t, D}), pos;
    if (H) memcpy(win.ptr(), att_hist, (size_t)H * D * 4);
    memcpy(win.ptr() + (size_t)H * D, a.ptr(), n * 4);
    RelPositionalEncoding pe; pe.get_pos_emb((size_t)H + t, pos);
    L.self_attn.forward(win, pos, b);
    for (size_t i = 0; i < n; i++) r.data[i] += b.data[(size_t)H * D + i];

    layer_norm(r, L.norm_conv_weight, L.norm_conv_bias, D, 1e-5f, a);
    memcpy(conv_new, a.ptr(), n * 4);
    TensorF cw({1, (size_t)Hc + t, D});
    if (Hc) memcpy(cw.ptr(), conv_hist, (size_t)Hc * D * 4);
    memcpy(cw.ptr() + (size_t)Hc * D, a.ptr(), n * 4);
    L.conv.forward(cw, b);
    for (size_t i = 0; i < n; i++) r.data[i] += b.data[(size_t)Hc * D + i];

    layer_norm(r, L.norm_ff2_weight, L.norm_ff2_bias, D, 1e-5f, a);
    L.ffn2.forward(a, b);
    for (size_t i = 0; i < n; i++) r.data[i] += 0.5f * b.data[i];
    TensorF o;
    layer_norm(r, L.norm_out_weight, L.norm_out_bias, D, 1e-5f, o);
    memcpy(y, o.ptr(), n * 4);
    return T;
}

// GreedyDecoder::decode over enc [T,1024] starting from a fresh decoder state
int ref_greedy(void* w, const float* enc, int T, int32_t* toks, int cap) {
    RNNTDecoder dec; RNNTJoint jn; dec.load_weights(*(ModelWeights*)w); jn.load_weights(*(ModelWeights*)w);
    GreedyDecoder g; g.init(&dec, &jn);
    TensorF e({1, (size_t)T, 1024}); memcpy(e.ptr(), enc, (size_t)T * 1024 * 4);
    std::vector<int> t = g.decode(e);
    for (int i = 0; i < (int)t.size() && i < cap; ++i) toks[i] = t[i];
    return (int)t.size();
}

// One joint evaluation: logits for (enc frame, decoder fed `token` from zero state)
int ref_joint_logits(void* w, const float* enc_frame, int token, float* logits) {
    RNNTDecoder dec; RNNTJoint jn; dec.load_weights(*(ModelWeights*)w); jn.load_weights(*(ModelWeights*)w);
    dec.init_state(1); TensorF d; dec.forward_step(token, d);
    TensorF e({1, 1024}); memcpy(e.ptr(), enc_frame, 1024 * 4);
    TensorF lg; jn.forward(e, d, lg); memcpy(logits, lg.ptr(), 1025 * 4); return 1025;
}

}  // extern "C"
