// nemo-ggml.h -- drop-in replacement for the reference's src/nemo-ggml.h on the streaming hot path.
//
// Same public names the CLI (src/transcribe_stream.cpp) and tests touch; no ggml behind it. Everything
// that was a ggml_tensor* / ggml_backend handle in the reference is an opaque pointer here (never
// dereferenced by callers) so that reference sources compile unmodified against this header.
// The compute lives in libnsb200.so (include/nsb200.h): hand-written sm_100a kernels, no CPU fallback.
#ifndef NEMO_GGML_H
#define NEMO_GGML_H

#include <algorithm>
#include <cstdint>
#include <map>
#include <string>
#include <vector>

// opaque stand-ins for the ggml types that appear in the reference's public structs
struct ggml_tensor; struct ggml_context; struct ggml_cgraph;
typedef struct nsb_opaque_gallocr* ggml_gallocr_t;
typedef struct nsb_opaque_backend* ggml_backend_t;
typedef struct nsb_opaque_buffer* ggml_backend_buffer_t;

struct timed_token;
struct nemo_preprocessor;
struct nsb_engine;

enum nemo_backend_type {          // src/nemo-ggml.h:26-31
    NEMO_BACKEND_CPU = 0,         // accepted, but there is no CPU path: runs on CUDA with a warning
    NEMO_BACKEND_CUDA = 1,
    NEMO_BACKEND_METAL = 2,       // not available: init fails
    NEMO_BACKEND_AUTO = 3,
};

struct nemo_hparams {             // src/nemo-ggml.h:37-49 (defaults = the reference's)
    int32_t n_mels = 128, d_model = 1024, n_heads = 8, d_head = 128, d_ff = 4096, n_layers = 24;
    int32_t kernel_size = 31, vocab_size = 1025, decoder_dim = 320, joint_dim = 640;
    float eps = 1e-5f;
};

struct char8 { char data[8]; };   // src/nemo-ggml.h:157-160

struct nemo_decoder {             // dims only (src/nemo-ggml.h:129-133); weights live in HBM
    static constexpr int NUM_LAYERS = 2, HIDDEN_SIZE = 640, EMBED_DIM = 640;
};

struct nemo_model {               // src/nemo-ggml.h:169-188 minus the ggml-typed weight structs
    nemo_hparams hparams;
    std::vector<char8> vocab;
    nemo_backend_type backend_type = NEMO_BACKEND_CUDA;
    std::string path;             // GGUF file the engines are created from
    int weight_type = 0;          // ggml type of the per-layer matrices
};

struct nemo_state {               // src/nemo-ggml.h:191-214 (batch path state; unused by streaming)
    static constexpr int HIDDEN_SIZE = 640, NUM_LAYERS = 2;
    std::vector<float> h, c; int prev_token = 1024; ggml_gallocr_t allocr = nullptr;
    nemo_state() : h(NUM_LAYERS * HIDDEN_SIZE, 0.0f), c(NUM_LAYERS * HIDDEN_SIZE, 0.0f) {}
    void reset() { std::fill(h.begin(), h.end(), 0.0f); std::fill(c.begin(), c.end(), 0.0f); prev_token = 1024; }
};

struct nemo_context {             // src/nemo-ggml.h:217-227
    nemo_model model;
    nemo_state state;
    struct nemo_preprocessor* preprocessor = nullptr;
    int n_threads = 4;            // kept for source compatibility; meaningless on the GPU path
    bool timestamp_words = false;
    // engines keyed by att_right_context (created on the first nemo_stream_init with that value)
    std::map<int, struct nsb_engine*> engines;
    int max_streams = 8;          // stream slots per engine; override with NSB_MAX_STREAMS
    struct nsb_engine* batch_engine = nullptr; int batch_rows = 0;   // nemo_transcribe_audio: created on first use (its batch workspace grows with the utterance)
};

struct nemo_context* nemo_init(const char* model_path);                                       // :231
struct nemo_context* nemo_init_with_backend(const char* model_path, nemo_backend_type backend); // :234
void nemo_free(struct nemo_context* ctx);                                                      // :236
const char* nemo_get_backend_name(struct nemo_context* ctx);                                   // :239
bool nemo_model_load(const std::string& path, nemo_model& model, nemo_backend_type backend = NEMO_BACKEND_AUTO);   // :242

struct timed_token {              // src/nemo-ggml.h:343-355
    int token_id; int64_t frame_idx;
    timed_token(int id = 0, int64_t frame = 0) : token_id(id), frame_idx(frame) {}
    float to_seconds(int frame_samples = 1280, int sample_rate = 16000) const { return (float)frame_idx * frame_samples / sample_rate; }
};

struct nemo_decoder_state {       // src/nemo-ggml.h:358-398
    int prev_token; std::vector<float> h, c; int32_t n_layers, hidden_size; int64_t frame_offset;
    nemo_decoder_state() : prev_token(-1), n_layers(0), hidden_size(0), frame_offset(0) {}
    void init(int32_t layers, int32_t hidden) { n_layers = layers; hidden_size = hidden; h.assign((size_t)layers * hidden, 0.0f);
        c.assign((size_t)layers * hidden, 0.0f); prev_token = -1; frame_offset = 0; }
    void reset() { std::fill(h.begin(), h.end(), 0.0f); std::fill(c.begin(), c.end(), 0.0f); prev_token = -1; frame_offset = 0; }
    void reset(int blank_token) { reset(); prev_token = blank_token; }
    bool is_initialized() const { return prev_token >= 0 && !h.empty() && !c.empty(); }
    float* h_layer(int l) { return h.data() + (size_t)l * hidden_size; }
    float* c_layer(int l) { return c.data() + (size_t)l * hidden_size; }
    const float* h_layer(int l) const { return h.data() + (size_t)l * hidden_size; }
    const float* c_layer(int l) const { return c.data() + (size_t)l * hidden_size; }
};

// src/nemo-ggml.h:330-339, src/nemo-ggml.cpp:1554-1600: the non-streaming batch path for one utterance (what src/transcribe.cpp
// calls), on nsb_transcribe_full (include/nsb200.h). nemo_encode_audio returns the tokens with their encoder frame
// (timed_token::frame_idx, for tokens_to_text(..., timestamp_words = true)); empty / "" on failure, like the reference.
std::vector<timed_token> nemo_encode_audio(struct nemo_context* ctx, std::vector<int16_t>& audio_data);
std::string nemo_transcribe_audio(struct nemo_context* ctx, std::vector<int16_t>& audio_data);

// src/nemo-ggml.cpp:1432-1458
std::string tokens_to_text(const std::vector<timed_token>& tokens, const std::vector<char8>& vocab, bool timestamp_words);

#endif  // NEMO_GGML_H
