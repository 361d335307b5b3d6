// nemo-stream.h -- drop-in replacement for the reference's src/nemo-stream.h (streaming API surface).
// Same names, argument meaning and error behaviour (nullptr / "" on failure, diagnostics on stderr);
// the work is done by libnsb200.so (include/nsb200.h). ggml graph members are gone: the per-stream
// caches, mel overlap and decoder state live in HBM inside the engine.
#ifndef NEMO_STREAM_H
#define NEMO_STREAM_H

#include <cstdint>
#include <string>
#include <vector>

#include "nemo-ggml.h"

enum class nemo_latency_mode { PURE_CAUSAL = 0, ULTRA_LOW = 1, LOW = 6, DEFAULT = 13 };   // src/nemo-stream.h:15-20

struct nemo_cache_config {        // src/nemo-stream.h:23-128 (field names and defaults are the API)
    int32_t att_left_context = 70, att_right_context = 0, cache_drop_size = 0;
    int32_t conv_kernel_size = 9, conv_cache_size = 8;
    int32_t d_model = 1024, n_layers = 24, n_heads = 8, d_head = 128;
    int32_t subsampling_factor = 8, n_mels = 128;
    int32_t sample_rate = 16000, hop_length = 160;
    int32_t decoder_hidden = 640, decoder_layers = 2, vocab_size = 1025, blank_token = 1024;
    int32_t drop_extra_pre_encoded = 2, last_channel_cache_size = 70, pre_encode_cache_size = 9, shift_mel_frames = 8;

    size_t get_chunk_mel_frames() const { return pre_encode_cache_size + subsampling_factor * (1 + att_right_context); }   // :65-72
    size_t get_shift_mel_frames() const { return subsampling_factor + subsampling_factor * (att_right_context - cache_drop_size); }   // :76-81
    int32_t get_chunk_samples() const { return (int32_t)get_chunk_mel_frames() * hop_length; }                           // :85-87
    int32_t get_latency_ms() const { return (int32_t)get_chunk_mel_frames() * hop_length * 1000 / sample_rate; }         // :90-92
    int32_t get_valid_out_len() const { return 1 + att_right_context; }                                                  // :98-100

    static nemo_cache_config with_latency(nemo_latency_mode mode) { nemo_cache_config c; c.att_right_context = (int32_t)mode; return c; }
    static nemo_cache_config default_config() { return with_latency(nemo_latency_mode::PURE_CAUSAL); }
    static nemo_cache_config pure_causal() { return with_latency(nemo_latency_mode::PURE_CAUSAL); }
    static nemo_cache_config ultra_low_latency() { return with_latency(nemo_latency_mode::ULTRA_LOW); }
    static nemo_cache_config low_latency() { return with_latency(nemo_latency_mode::LOW); }
    static nemo_cache_config balanced() { return with_latency(nemo_latency_mode::DEFAULT); }
};

struct nemo_stream_context {      // src/nemo-stream.h:176-253, public (non-ggml) fields kept
    struct nemo_context* nctx = nullptr;       // borrowed; must outlive the stream
    nemo_cache_config config;
    nemo_decoder_state decoder_state;           // host mirror is NOT kept in sync (state lives in HBM)
    std::vector<float> mel_buffer;              // unused (mel overlap lives in HBM)
    std::vector<int> tokens;
    std::string transcript;
    double total_audio_seconds = 0, total_compute_seconds = 0;
    double encoder_seconds = 0, decoder_seconds = 0, transfer_seconds = 0;
    int total_decode_iterations = 0;
    int cache_valid_len = 0;
    int total_chunks_processed = 0;
    // engine binding
    struct nsb_engine* engine = nullptr; int stream_id = -1;

    size_t overlap_mel_frames() const { return config.pre_encode_cache_size; }
    size_t shift_mel_frames() const { return config.get_shift_mel_frames(); }
    double rtf() const { return total_audio_seconds > 0 ? total_compute_seconds / total_audio_seconds : 0; }
};

struct nemo_stream_context* nemo_stream_init(struct nemo_context* ctx, const nemo_cache_config* config = nullptr);   // :262-265
std::string nemo_stream_process_incremental(struct nemo_stream_context* sctx, const int16_t* audio, int n_samples);  // :282-286
std::string nemo_stream_finalize(struct nemo_stream_context* sctx);                                                  // :290-292
std::string nemo_stream_get_transcript(struct nemo_stream_context* sctx);                                            // :295-297
const std::vector<int>& nemo_stream_get_tokens(struct nemo_stream_context* sctx);                                    // :300-302
void nemo_stream_reset(struct nemo_stream_context* sctx);                                                            // :305-307
void nemo_stream_free(struct nemo_stream_context* sctx);                                                             // :310-312

#endif  // NEMO_STREAM_H
