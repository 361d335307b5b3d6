// nemo_shim.cpp -- host-side mirror of the reference's C++ API (include/nemo-ggml.h, include/nemo-stream.h,
// include/preprocessor.h) on top of the C ABI (include/nsb200.h). Pure host C++: no CUDA, no ggml.
// Error behaviour follows the reference: nullptr / false / "" + a line on stderr, no exceptions.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "nemo-stream.h"
#include "nsb200.h"
#include "preprocessor.h"

// ------------------------------------------------------------------------------------------
// preprocessor.h
// ------------------------------------------------------------------------------------------
struct nemo_preprocessor {
    std::vector<float> fb, window;
    nsb_engine* engine = nullptr;        // bound by nemo_stream_init (first engine of the context)
    std::vector<int16_t> tail;           // stand-alone API: audio seen so far (stateful like the reference)
    size_t frames_emitted = 0;
};

struct nemo_preprocessor* nemo_preprocessor_init(const char*, const char*) {
    fprintf(stderr, "nemo_preprocessor_init(path): not supported (the reference variant indexes a 400-tap window as 512); "
                    "use nemo_preprocessor_init_from_data\n");
    return nullptr;
}
struct nemo_preprocessor* nemo_preprocessor_init_from_data(const float* fb, size_t fb_size, const float* win, size_t win_size) {
    if (win_size != 400) { fprintf(stderr, "Window size mismatch: got %zu, expected 400\n", win_size); return nullptr; }
    if (fb_size != 128 * 257) { fprintf(stderr, "Filterbank size mismatch: got %zu, expected %d\n", fb_size, 128 * 257); return nullptr; }
    nemo_preprocessor* pp = new nemo_preprocessor();
    pp->fb.assign(fb, fb + fb_size); pp->window.assign(win, win + win_size);
    return pp;
}
void nemo_preprocessor_free(struct nemo_preprocessor* pp) { delete pp; }
size_t nemo_preprocessor_get_n_frames(struct nemo_preprocessor*, size_t n_samples) {
    if (n_samples == 0) return 0;
    return 1 + (n_samples + 512 - 512) / 160;                      // src/preprocessor.cpp:313-318
}
// Stateful semantics of src/preprocessor.cpp:330-395 reproduced by re-running the batched kernel over the
// retained audio and returning only frames not handed out yet (debug/compat API; the hot path never calls it).
size_t nemo_preprocessor_process(struct nemo_preprocessor* pp, const int16_t* audio, size_t n, std::vector<float>& mel_out) {
    mel_out.clear();
    if (!pp || !audio || n == 0) return 0;
    if (!pp->engine) { fprintf(stderr, "nemo_preprocessor_process: no engine bound (create a stream first)\n"); return 0; }
    pp->tail.insert(pp->tail.end(), audio, audio + n);
    const size_t avail = 256 + pp->tail.size();
    const size_t total = avail < 512 ? 0 : (avail - 512 + 160) / 160;
    if (total <= pp->frames_emitted) return 0;
    std::vector<float> all(total * 128);
    int nf = nsb_op_logmel(pp->engine, pp->tail.data(), 1, (int)pp->tail.size(), all.data(), all.size());
    if (nf < 0) { fprintf(stderr, "nemo_preprocessor_process: %s\n", nsb_last_error()); return 0; }
    mel_out.assign(all.begin() + pp->frames_emitted * 128, all.begin() + (size_t)nf * 128);
    const size_t out = (size_t)nf - pp->frames_emitted;
    pp->frames_emitted = (size_t)nf;
    return out;
}

// ------------------------------------------------------------------------------------------
// nemo-ggml.h
// ------------------------------------------------------------------------------------------
bool nemo_model_load(const std::string& path, nemo_model& model, nemo_backend_type backend) {
    if (backend == NEMO_BACKEND_METAL) { fprintf(stderr, "%s: Metal backend not available\n", __func__); return false; }
    if (backend == NEMO_BACKEND_CPU) fprintf(stderr, "%s: warning: no CPU path in this build; running on CUDA (B200)\n", __func__);
    nsb_model_info* info = new nsb_model_info();
    if (nsb_gguf_probe(path.c_str(), info) != NSB_OK) {
        fprintf(stderr, "%s: failed to open GGUF file: %s\n", __func__, nsb_last_error());
        delete info; return false;
    }
    nemo_hparams& h = model.hparams;
    h.n_mels = info->n_mels; h.d_model = info->d_model; h.n_heads = info->n_heads; h.d_head = info->d_head; h.d_ff = info->d_ff;
    h.n_layers = info->n_layers; h.kernel_size = info->kernel_size; h.vocab_size = info->vocab_size;
    h.decoder_dim = info->decoder_dim; h.joint_dim = info->joint_dim;
    model.vocab.assign((size_t)h.vocab_size, char8{{0}});
    memcpy(model.vocab.data(), info->vocab, std::min(sizeof(info->vocab), model.vocab.size() * 8));
    model.backend_type = NEMO_BACKEND_CUDA; model.path = path; model.weight_type = info->weight_type;
    delete info;
    return true;
}

struct nemo_context* nemo_init(const char* model_path) { return nemo_init_with_backend(model_path, NEMO_BACKEND_AUTO); }

struct nemo_context* nemo_init_with_backend(const char* model_path, nemo_backend_type backend) {
    if (!model_path) return nullptr;
    nemo_context* ctx = new nemo_context();
    if (!nemo_model_load(model_path, ctx->model, backend)) { delete ctx; return nullptr; }
    if (const char* e = getenv("NSB_MAX_STREAMS")) ctx->max_streams = std::max(1, atoi(e));
    ctx->preprocessor = new nemo_preprocessor();
    return ctx;
}
const char* nemo_get_backend_name(struct nemo_context* ctx) { return ctx ? "CUDA" : "unknown"; }
void nemo_free(struct nemo_context* ctx) {
    if (!ctx) return;
    for (auto& kv : ctx->engines) nsb_engine_destroy(kv.second);
    if (ctx->batch_engine) nsb_engine_destroy(ctx->batch_engine);
    nemo_preprocessor_free(ctx->preprocessor);
    delete ctx;
}

std::string tokens_to_text(const std::vector<timed_token>& tokens, const std::vector<char8>& vocab, bool timestamp_words) {
    std::string r;
    for (const timed_token& t : tokens) {
        if (t.token_id < 0 || t.token_id >= (int)vocab.size()) continue;
        char piece[9]; memcpy(piece, vocab[t.token_id].data, 8); piece[8] = 0;
        if (strncmp(piece, "\xe2\x96\x81", 3) == 0) {
            r += ' ';
            if (timestamp_words) { char b[32]; snprintf(b, sizeof(b), "{%.2f}", t.to_seconds()); r += b; }
            r += piece + 3;
        } else r += piece;
    }
    return r;
}

// nemo_encode_audio / nemo_transcribe_audio (src/nemo-ggml.cpp:1554-1600): mel of the whole utterance -> nemo_encode -> timed tokens
// (-> tokens_to_text). Here one call into the engine's batch path, whose workspace grows with the utterance: one engine serves every
// length up to the reference's own limit of 2048 encoder frames.
std::vector<timed_token> nemo_encode_audio(struct nemo_context* ctx, std::vector<int16_t>& audio) {
    std::vector<timed_token> toks;
    if (!ctx || audio.empty()) return toks;
    const long long avail = 256 + (long long)audio.size();
    const long long mel = avail < 512 ? 0 : (avail - 512 + 160) / 160;            // preprocessor.cpp:320-328
    if (mel <= 0) return toks;
    const int frames = (int)(((mel / 2 + 1) / 2 + 1) / 2 + 1);                     // three stride-2 convs, (2, 1) padding (nemo-ggml.cpp:828-836)
    if (!ctx->batch_engine) {
        nsb_engine_config ec; nsb_default_config(&ec);
        ec.att_right_context = 13; ec.max_streams = 1;
        if (const char* e = getenv("NSB_COMPUTE")) ec.compute = atoi(e);
        if (const char* e = getenv("NSB_KV_DTYPE")) ec.kv_dtype = atoi(e);
        if (const char* e = getenv("NSB_DEVICE")) ec.device = atoi(e);
        if (nsb_engine_create(ctx->model.path.c_str(), &ec, &ctx->batch_engine) != NSB_OK) {
            fprintf(stderr, "[ERROR] Failed to create engine: %s\n", nsb_last_error());
            ctx->batch_engine = nullptr; return toks;
        }
    }
    std::vector<int32_t> ids((size_t)frames * 10 + 1), frm((size_t)frames * 10 + 1);   // <= 10 symbols per frame (nemo-ggml.cpp:1132)
    int nf = 0;
    const int n = nsb_transcribe_full(ctx->batch_engine, audio.data(), (int)audio.size(), ids.data(), frm.data(), (int)ids.size(), &nf, nullptr, 0);
    if (n < 0) { fprintf(stderr, "[ERROR] nemo_encode_audio: %s\n", nsb_last_error()); return toks; }
    for (int i = 0; i < n && i < (int)ids.size(); ++i) toks.push_back({ids[i], frm[i]});   // frame index -> timed_token::to_seconds (nemo-ggml.cpp:1240)
    return toks;
}
std::string nemo_transcribe_audio(struct nemo_context* ctx, std::vector<int16_t>& audio) {
    if (!ctx) return "";
    const std::vector<timed_token> toks = nemo_encode_audio(ctx, audio);
    return toks.empty() ? std::string() : tokens_to_text(toks, ctx->model.vocab, false);
}

// ------------------------------------------------------------------------------------------
// nemo-stream.h
// ------------------------------------------------------------------------------------------
struct nemo_stream_context* nemo_stream_init(struct nemo_context* ctx, const nemo_cache_config* config) {
    if (!ctx) return nullptr;
    nemo_cache_config cfg;
    if (config) cfg = *config;
    else {                                                           // src/nemo-stream.cpp:680-692
        cfg.d_model = ctx->model.hparams.d_model; cfg.n_layers = ctx->model.hparams.n_layers; cfg.n_heads = ctx->model.hparams.n_heads;
        cfg.d_head = ctx->model.hparams.d_head; cfg.vocab_size = ctx->model.hparams.vocab_size; cfg.blank_token = cfg.vocab_size - 1;
    }
    if (cfg.att_left_context != 70 || cfg.conv_kernel_size != 9 || cfg.pre_encode_cache_size != 9 || cfg.drop_extra_pre_encoded != 2 ||
        cfg.subsampling_factor != 8 || cfg.cache_drop_size != 0) {
        fprintf(stderr, "[ERROR] nemo_stream_init: only the model's streaming geometry (L=70, k=9, overlap 9, drop 2) is built\n");
        return nullptr;
    }
    nsb_engine*& eng = ctx->engines[cfg.att_right_context];
    if (!eng) {
        nsb_engine_config ec; nsb_default_config(&ec);
        ec.att_right_context = cfg.att_right_context; ec.max_streams = ctx->max_streams;
        if (const char* e = getenv("NSB_COMPUTE")) ec.compute = atoi(e);
        if (const char* e = getenv("NSB_KV_DTYPE")) ec.kv_dtype = atoi(e);
        if (const char* e = getenv("NSB_DEVICE")) ec.device = atoi(e);
        if (nsb_engine_create(ctx->model.path.c_str(), &ec, &eng) != NSB_OK) {
            fprintf(stderr, "[ERROR] Failed to create engine: %s\n", nsb_last_error());
            ctx->engines.erase(cfg.att_right_context);
            return nullptr;
        }
        if (ctx->preprocessor && !ctx->preprocessor->engine) ctx->preprocessor->engine = eng;
    }
    const int sid = nsb_stream_open(eng);
    if (sid < 0) { fprintf(stderr, "[ERROR] Failed to open stream: %s\n", nsb_last_error()); return nullptr; }
    nemo_stream_context* s = new nemo_stream_context();
    s->nctx = ctx; s->config = cfg; s->engine = eng; s->stream_id = sid;
    s->decoder_state.init(cfg.decoder_layers, cfg.decoder_hidden);
    s->decoder_state.prev_token = cfg.blank_token;
    return s;
}

std::string nemo_stream_process_incremental(struct nemo_stream_context* s, const int16_t* audio, int n_samples) {
    if (!s || !audio || n_samples <= 0) return "";
    const auto t0 = std::chrono::high_resolution_clock::now();
    s->total_audio_seconds += (double)n_samples / s->config.sample_rate;
    if (nsb_stream_push_pcm(s->engine, s->stream_id, audio, n_samples) != NSB_OK) { fprintf(stderr, "[ERROR] %s\n", nsb_last_error()); return ""; }
    // run every chunk this stream has buffered (the reference loops "while total_mels >= chunk": nemo-stream.cpp:1102-1127)
    while (nsb_stream_ready(s->engine, s->stream_id) == 1) {
        const int n = nsb_engine_step(s->engine);
        if (n < 0) { fprintf(stderr, "[ERROR] %s\n", nsb_last_error()); return ""; }
        if (n == 0) break;
    }
    std::vector<timed_token> fresh; int32_t buf[256];
    for (;;) {
        const int n = nsb_stream_pop_tokens(s->engine, s->stream_id, buf, 256);
        if (n <= 0) break;
        for (int i = 0; i < n; ++i) { fresh.push_back({buf[i], 0}); s->tokens.push_back(buf[i]); }
    }
    s->total_chunks_processed = nsb_stream_chunks(s->engine, s->stream_id);
    s->cache_valid_len = std::min(70, s->total_chunks_processed * s->config.get_valid_out_len());
    nsb_stats st; nsb_engine_get_stats(s->engine, &st); s->encoder_seconds = st.device_ms * 1e-3;
    std::string text = fresh.empty() ? std::string() : tokens_to_text(fresh, s->nctx->model.vocab, false);
    s->transcript += text;
    s->total_compute_seconds += std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
    return text;
}

// The reference does NOT flush buffered audio here (nemo-stream.cpp:1137-1172): it returns the whole transcript.
std::string nemo_stream_finalize(struct nemo_stream_context* s) { return s ? s->transcript : std::string(); }
std::string nemo_stream_get_transcript(struct nemo_stream_context* s) { return s ? s->transcript : std::string(); }
const std::vector<int>& nemo_stream_get_tokens(struct nemo_stream_context* s) { static std::vector<int> empty; return s ? s->tokens : empty; }

void nemo_stream_reset(struct nemo_stream_context* s) {
    if (!s) return;
    nsb_stream_reset(s->engine, s->stream_id);       // also re-zeroes the device caches (the reference's reset leaves them stale)
    s->decoder_state.reset(); s->decoder_state.prev_token = s->config.blank_token;
    s->tokens.clear(); s->transcript.clear();
    s->total_audio_seconds = s->total_compute_seconds = s->encoder_seconds = s->decoder_seconds = s->transfer_seconds = 0;
    s->total_decode_iterations = 0; s->cache_valid_len = 0; s->total_chunks_processed = 0;
}
void nemo_stream_free(struct nemo_stream_context* s) {
    if (!s) return;
    if (s->engine && s->stream_id >= 0) nsb_stream_close(s->engine, s->stream_id);
    delete s;
}
