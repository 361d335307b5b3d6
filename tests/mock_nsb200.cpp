// tests/mock_nsb200.cpp -- TEST DOUBLE of the C ABI (include/nsb200.h) for the CPU test of nemotron-asr-serve's host loop: no model,
// no arithmetic. Chunk gating is the engine's own host code (csrc/host_stream.h); every processed chunk "decodes" to one token =
// its chunk index, so the test can see order, loss and duplication. Linked ONLY by tests/test_serve_cli.py into a scratch binary.
#include <cstdio>
#include <cstring>
#include <deque>
#include <string>
#include <utility>
#include <vector>

#include "nsb200.h"
#include "host_stream.h"

struct nsb_engine {
    int T = 1, max_streams = 1, compute = NSB_COMPUTE_F32;
    std::vector<nsb::HostStream> hs;
    std::deque<std::vector<std::pair<int, int>>> inflight;     // (stream, token) per step
    nsb_stats st{};
};
static thread_local std::string g_err;
static int fail(int code, const char* m) { g_err = m; return code; }

extern "C" {
const char* nsb_last_error(void) { return g_err.c_str(); }
int nsb_gguf_probe(const char* p, nsb_model_info* info) {
    if (!p || !info) return fail(NSB_ERR_ARG, "null argument");
    if (strstr(p, "missing")) return fail(NSB_ERR_IO, "gguf: cannot open (mock)");
    memset(info, 0, sizeof(*info)); info->n_layers = 2; info->vocab_size = 1025; info->d_model = 1024; info->n_heads = 8; info->d_head = 128;
    for (int i = 0; i < 1025; ++i) snprintf(info->vocab + 8 * i, 8, "%d,", i);                 // piece of token i = "i,"
    return NSB_OK;
}
void nsb_default_config(nsb_engine_config* c) { memset(c, 0, sizeof(*c)); c->max_streams = 1; c->use_cuda_graph = 1; }
int nsb_engine_create(const char*, const nsb_engine_config* cfg, nsb_engine** out) {
    auto* e = new nsb_engine(); e->T = 1 + cfg->att_right_context; e->max_streams = cfg->max_streams; e->hs.resize((size_t)cfg->max_streams);
    e->compute = cfg->compute ? cfg->compute : NSB_COMPUTE_F32; *out = e; return NSB_OK;
}
void nsb_engine_destroy(nsb_engine* e) { delete e; }
int nsb_engine_chunk_samples(const nsb_engine* e) { return (9 + 8 * e->T) * 160; }
int nsb_engine_shift_samples(const nsb_engine* e) { return 8 * e->T * 160; }
int nsb_engine_compute(const nsb_engine* e) { return e->compute; }
int nsb_engine_step_end(nsb_engine* e);
static void collect(nsb_engine* e) { while (!e->inflight.empty()) nsb_engine_step_end(e); }
int nsb_stream_open(nsb_engine* e) {
    collect(e);
    for (int s = 0; s < e->max_streams; ++s) if (!e->hs[s].open) { nsb::hs_clear(e->hs[s]); e->hs[s].open = true; return s; }
    return fail(NSB_ERR_STATE, "no free stream slot");
}
int nsb_stream_close(nsb_engine* e, int s) { collect(e); if (s < 0 || s >= e->max_streams || !e->hs[s].open) return fail(NSB_ERR_ARG, "bad stream id"); e->hs[s].open = false; return NSB_OK; }
int nsb_stream_reset(nsb_engine* e, int s) { collect(e); if (s < 0 || s >= e->max_streams || !e->hs[s].open) return fail(NSB_ERR_ARG, "bad stream id"); nsb::hs_clear(e->hs[s]); return NSB_OK; }
int nsb_stream_push_pcm(nsb_engine* e, int s, const int16_t* pcm, int n) {
    if (s < 0 || s >= e->max_streams || !e->hs[s].open) return fail(NSB_ERR_ARG, "bad stream id");
    if (pcm && n > 0) nsb::hs_push(e->hs[s], pcm, n);
    return NSB_OK;
}
int nsb_engine_step_begin(nsb_engine* e) {
    if (e->inflight.size() == 3) return fail(NSB_ERR_STATE, "step_begin: three steps are already in flight");
    std::vector<std::pair<int, int>> batch;
    std::vector<int16_t> row((size_t)nsb::hs_row_len(e->T));
    for (int s = 0; s < e->max_streams; ++s)
        if (nsb::hs_ready(e->hs[s], e->T)) {
            nsb::hs_stage_row(e->hs[s], e->T, (int)row.size(), row.data());           // exercises the buffer bounds like the engine does
            batch.push_back({s, (int)(e->hs[s].chunk_idx % 1024)});
            nsb::hs_launched(e->hs[s], e->T);
        }
    if (batch.empty()) return 0;
    const int B = (int)batch.size();
    e->inflight.push_back(std::move(batch));
    return B;
}
int nsb_engine_step_end(nsb_engine* e) {
    if (e->inflight.empty()) return 0;
    auto batch = std::move(e->inflight.front()); e->inflight.pop_front();
    for (auto& p : batch) { e->hs[p.first].tokens.push_back(p.second); e->hs[p.first].chunks_done += 1; }
    e->st.steps += 1; e->st.chunks += (long long)batch.size(); e->st.kernel_launches += 1; e->st.device_ms += 0.5; e->st.last_step_ms = 0.5;
    return (int)batch.size();
}
int nsb_engine_step(nsb_engine* e) { collect(e); const int B = nsb_engine_step_begin(e); if (B > 0) nsb_engine_step_end(e); return B; }
int nsb_stream_pop_tokens(nsb_engine* e, int s, int32_t* out, int cap) {
    int n = 0; auto& q = e->hs[s].tokens; while (n < cap && !q.empty()) { out[n++] = q.front(); q.pop_front(); } return n;
}
int nsb_pop_tokens_batch(nsb_engine* e, int n, const int32_t* ids, int32_t* out, int cap, int32_t* counts) {
    int total = 0; for (int i = 0; i < n; ++i) { counts[i] = nsb_stream_pop_tokens(e, ids[i], out + (size_t)i * cap, cap); total += counts[i]; } return total;
}
int nsb_stream_chunks(const nsb_engine* e, int s) { return (int)e->hs[s].chunks_done; }
int nsb_detokenize(const nsb_engine*, const int32_t* t, int n, char* out, int cap) {
    std::string r; for (int i = 0; i < n; ++i) r += "<" + std::to_string(t[i]) + ">";
    if ((int)r.size() + 1 > cap) return -(int)r.size() - 1;
    memcpy(out, r.c_str(), r.size() + 1); return (int)r.size();
}
void nsb_engine_get_stats(const nsb_engine* e, nsb_stats* out) { *out = e->st; }
int nsb_stream_ready(const nsb_engine* e, int s) { return s >= 0 && s < e->max_streams && nsb::hs_ready(e->hs[s], e->T) ? 1 : 0; }
// referenced by the drop-in shim (csrc/nemo_shim.cpp) but not part of what the double models
int nsb_op_logmel(nsb_engine*, const int16_t*, int, int, float*, size_t) { return fail(NSB_ERR_STATE, "mock: no log-mel"); }
int nsb_transcribe_full(nsb_engine*, const int16_t*, int, int32_t*, int32_t*, int, int*, float*, size_t) { return fail(NSB_ERR_STATE, "mock: no batch path"); }
}
