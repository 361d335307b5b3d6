// serve_main.cpp -- nemotron-asr-serve: many PCM streams through the batched engine, sharded over the GPUs of one box.
//
// The reference CLI (src/transcribe_stream.cpp) feeds ONE stream through nemo_stream_process_incremental; the engine under
// the drop-in API is batched, so this is the host program for the multi-stream case: plain C++ over the
// C ABI (include/nsb200.h), no torch, no NCCL. Stream s lives on GPU s mod G for its whole life (its caches live there),
// one engine + one host thread per GPU, transcripts gathered on the host in input order.
//
//   nemotron-asr-serve model.gguf [options] a.pcm b.pcm ...        raw 16 kHz s16le mono files, one stream each
//     --list FILE            more input paths, one per line
//     --synthetic N SECONDS  N generated streams (sines + noise, seeded per stream) instead of / besides files
//     --right-context R      0 | 1 | 6 | 13 (nemo_cache_config::att_right_context; chunk = 80 ms x (R+1)); default 13 as README "70 13"
//     --compute auto|f32|f16|bf16|q8_0     --kv f32|f16|bf16
//     --gpus G | --devices 0,2,...          default: device 0
//     --max-streams N        stream slots per GPU (default 256); more inputs than slots run in waves
//     --realtime             pace every stream at audio rate (one chunk shift per tick), one step in flight; reports the
//                            per-chunk latency (last sample handed over -> token ids back on the host)
//     --flush                zero-pad each stream's tail up to the next chunk boundary so that the last partial chunk is
//                            decoded (the reference drops it: transcribe_stream.cpp:143-166 never pads, nemo-stream.cpp:1102)
//     --warmup               before a wave starts, run one chunk of silence through all its streams and reset them: the step's CUDA
//                            graph for that batch size is captured outside the measured run (off by default)
//     --tokens               print token ids after the text
// stdout: "<index>\t<name>\t<transcript>" per stream, input order. stderr: per-GPU and total statistics.
// Exit code 1 on usage / load / open failure (transcribe_stream.cpp:53-56,102-105,131-137), 2 on an engine error mid-run.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <thread>
#include <vector>

#include "../../include/nsb200.h"

namespace {

using Clock = std::chrono::steady_clock;
double seconds_since(Clock::time_point t0) { return std::chrono::duration<double>(Clock::now() - t0).count(); }

struct Input {
    std::string name;
    std::vector<int16_t> pcm;
};
struct Result {
    std::vector<int32_t> tokens;
    std::string text;
    int chunks = 0;
};
struct Options {
    std::string model;
    int right_context = 13, compute = NSB_COMPUTE_AUTO, kv = -1, max_streams = 256;
    std::vector<int> devices{0};
    bool realtime = false, flush = false, print_tokens = false, warmup = false;
};
struct WorkerStats {
    int device = 0, streams = 0, waves = 0;
    long long chunks = 0, steps = 0, launches = 0;
    double device_ms = 0, wall_s = 0, audio_s = 0, load_s = 0;
    std::vector<double> latency_ms;   // --realtime only: one sample per step
    std::string error;
};

int usage(const char* argv0) {
    fprintf(stderr,
            "Usage: %s model.gguf [--right-context 0|1|6|13] [--compute auto|f32|f16|bf16|q8_0] [--kv f32|f16|bf16]\n"
            "          [--gpus G | --devices 0,1,..] [--max-streams N] [--realtime] [--flush] [--warmup] [--tokens]\n"
            "          [--list FILE] [--synthetic N SECONDS] [audio.pcm ...]\n"
            "  audio: raw 16 kHz s16le mono, one stream per file\n",
            argv0);
    return 1;
}

bool read_pcm(const std::string& path, std::vector<int16_t>& out) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) return false;
    const std::streamoff bytes = f.tellg();
    f.seekg(0);
    out.resize((size_t)(bytes / 2));
    if (!out.empty()) f.read(reinterpret_cast<char*>(out.data()), (std::streamsize)out.size() * 2);
    return (bool)f || f.eof();
}

// deterministic load-test audio: three sines under a slow amplitude envelope + a little white noise, +-0.5 full scale
std::vector<int16_t> synthetic_pcm(uint32_t seed, double seconds) {
    uint64_t s = 0x9E3779B97F4A7C15ull * (seed + 1);
    auto rnd = [&s] { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (double)(s >> 11) / 9007199254740992.0; };
    double f[3], ph[3];
    for (int i = 0; i < 3; ++i) { f[i] = 100.0 + 2900.0 * rnd(); ph[i] = 6.283185307179586 * rnd(); }
    const double am = 0.5 + 2.5 * rnd();
    const size_t n = (size_t)(seconds * 16000.0);
    std::vector<int16_t> out(n);
    for (size_t i = 0; i < n; ++i) {
        const double t = (double)i / 16000.0;
        double v = 0;
        for (int k = 0; k < 3; ++k) v += std::sin(6.283185307179586 * f[k] * t + ph[k]);
        v = v / 3.0 * (0.6 + 0.4 * std::sin(6.283185307179586 * am * t)) + 0.05 * (2.0 * rnd() - 1.0);
        out[i] = (int16_t)std::lrint(std::max(-1.0, std::min(1.0, v)) * 0.5 * 32767.0);
    }
    return out;
}

int parse_compute(const std::string& v) {
    if (v == "auto") return NSB_COMPUTE_AUTO;
    if (v == "f32") return NSB_COMPUTE_F32;
    if (v == "f16") return NSB_COMPUTE_F16;
    if (v == "bf16") return NSB_COMPUTE_BF16;
    if (v == "q8_0") return NSB_COMPUTE_Q8_0;
    return -1;
}
int parse_kv(const std::string& v) {
    if (v == "f32") return NSB_KV_F32;
    if (v == "f16") return NSB_KV_F16;
    if (v == "bf16") return NSB_KV_BF16;
    return -1;
}

// One GPU: its engine, its share of the inputs (indices `mine` into `inputs`), results written to the same indices.
void run_worker(const Options& opt, int device, const std::vector<Input>& inputs, const std::vector<int>& mine,
                std::vector<Result>& results, WorkerStats& ws) {
    ws.device = device;
    ws.streams = (int)mine.size();
    if (mine.empty()) return;
    const auto t_load = Clock::now();
    nsb_engine_config cfg;
    nsb_default_config(&cfg);
    cfg.device = device;
    cfg.compute = opt.compute;
    cfg.att_right_context = opt.right_context;
    cfg.max_streams = std::min<int>(opt.max_streams, (int)mine.size());
    cfg.kv_dtype = opt.kv;
    nsb_engine* e = nullptr;
    if (nsb_engine_create(opt.model.c_str(), &cfg, &e) != NSB_OK) { ws.error = std::string("Failed to load model: ") + nsb_last_error(); return; }
    ws.load_s = seconds_since(t_load);
    const int chunk = nsb_engine_chunk_samples(e), shift = nsb_engine_shift_samples(e), cap = cfg.max_streams;
    const int tok_cap = 10 * (opt.right_context + 1) * 4 + 16;      // <= 10 symbols per encoder frame (nemo-stream.cpp:797), a few steps deep
    std::vector<int32_t> tok((size_t)cap * tok_cap), cnt(cap), ids(cap);
    auto fail = [&](const char* what) { ws.error = std::string(what) + ": " + nsb_last_error(); };

    const auto t_run = Clock::now();
    for (size_t w0 = 0; w0 < mine.size() && ws.error.empty(); w0 += (size_t)cap) {         // one wave = up to `cap` streams side by side
        const int n = (int)std::min<size_t>((size_t)cap, mine.size() - w0);
        ++ws.waves;
        std::vector<size_t> pos(n, 0), len(n);
        std::vector<int> pad(n, 0);
        for (int i = 0; i < n; ++i) {
            ids[i] = nsb_stream_open(e);
            if (ids[i] < 0) { fail("nsb_stream_open"); break; }
            const size_t have = inputs[mine[w0 + i]].pcm.size();
            len[i] = have;
            if (opt.flush && have > 0) {
                // mel frame t is centred on raw sample 160 t (256-sample left pad, preprocessor.cpp:220-221), so the frames that look at
                // real audio are t <= (have - 1) / 160; a chunk consumes shift / 160 of them, and chunk number C is complete once
                // 160 (8T C - 1) + 256 = shift C + 96 samples have arrived (nemo-stream.cpp:1094-1102 gate on the frame count)
                const long long frames = ((long long)have - 1) / 160 + 1, per = shift / 160;
                const long long need = (long long)shift * ((frames + per - 1) / per) + 96;
                pad[i] = (int)std::max<long long>(0, need - (long long)have);
            }
        }
        if (!ws.error.empty()) break;
        static const std::vector<int16_t> zeros(1 << 16, 0);
        if (opt.warmup) {                                                         // capture the graph of an n-stream step on silence, then start over
            for (int i = 0; i < n && ws.error.empty(); ++i)
                for (int left = chunk; left > 0 && ws.error.empty(); left -= (int)zeros.size())
                    if (nsb_stream_push_pcm(e, ids[i], zeros.data(), std::min<int>(left, (int)zeros.size())) != NSB_OK) fail("nsb_stream_push_pcm");
            if (ws.error.empty() && nsb_engine_step(e) < 0) fail("nsb_engine_step");
            for (int i = 0; i < n && ws.error.empty(); ++i) if (nsb_stream_reset(e, ids[i]) != NSB_OK) fail("nsb_stream_reset");   // caches, decoder state, queued tokens
            if (!ws.error.empty()) break;
        }
        nsb_stats st0;
        nsb_engine_get_stats(e, &st0);                                            // statistics of the wave exclude the warm-up step
        // hand stream i its next `want` samples (real audio first, then the flush padding); returns samples handed over
        auto feed_one = [&](int i, int want) -> int {
            const std::vector<int16_t>& pcm = inputs[mine[w0 + i]].pcm;
            int given = 0;
            if (pos[i] < len[i]) {
                const int k = (int)std::min<size_t>((size_t)want, len[i] - pos[i]);
                if (nsb_stream_push_pcm(e, ids[i], pcm.data() + pos[i], k) != NSB_OK) { fail("nsb_stream_push_pcm"); return -1; }
                pos[i] += (size_t)k; given += k;
            }
            while (given < want && pad[i] > 0) {
                const int k = std::min<int>(std::min<int>(want - given, pad[i]), (int)zeros.size());
                if (nsb_stream_push_pcm(e, ids[i], zeros.data(), k) != NSB_OK) { fail("nsb_stream_push_pcm"); return -1; }
                pad[i] -= k; given += k;
            }
            return given;
        };
        auto feed_all = [&](int want) -> long long {
            long long total = 0;
            for (int i = 0; i < n; ++i) { const int g = feed_one(i, want); if (g < 0) return -1; total += g; }
            return total;
        };
        auto pop_all = [&]() -> bool {
            if (nsb_pop_tokens_batch(e, n, ids.data(), tok.data(), tok_cap, cnt.data()) < 0) { fail("nsb_pop_tokens_batch"); return false; }
            for (int i = 0; i < n; ++i) {
                Result& r = results[mine[w0 + i]];
                r.tokens.insert(r.tokens.end(), tok.begin() + (size_t)i * tok_cap, tok.begin() + (size_t)i * tok_cap + cnt[i]);
            }
            return true;
        };

        if (opt.realtime) {
            // live feeds: every tick each stream delivers one chunk shift of audio; the step runs as soon as it is in, and the
            // latency of that chunk = hand-over of its last sample -> its token ids on the host
            const auto t0 = Clock::now();
            long long tick = 0;
            if (feed_all(chunk - shift) < 0) break;
            for (;;) {
                const long long pushed = feed_all(shift);
                if (pushed < 0) break;
                const auto t_in = Clock::now();
                const int adv = nsb_engine_step(e);
                if (adv < 0) { fail("nsb_engine_step"); break; }
                if (!pop_all()) break;
                if (adv > 0) ws.latency_ms.push_back(1e3 * seconds_since(t_in));
                if (pushed == 0 && adv == 0) break;
                ++tick;
                std::this_thread::sleep_until(t0 + std::chrono::duration_cast<Clock::duration>(std::chrono::duration<double>(tick * shift / 16000.0)));
            }
        } else {
            // throughput: up to three steps in flight -- while the device runs step i the host hands over the next shifts of every
            // stream and stages + enqueues steps i+1, i+2 (nsb200.h: begin, begin, begin, end, begin, end, ...)
            constexpr int DEPTH = 3;
            int inflight = 0;
            if (feed_all(chunk) < 0) break;
            for (;;) {
                int launched = 0;
                if (inflight < DEPTH) {
                    launched = nsb_engine_step_begin(e);
                    if (launched < 0) { fail("nsb_engine_step_begin"); break; }
                    if (launched > 0) ++inflight;
                }
                const long long pushed = feed_all(shift);
                if (pushed < 0) break;
                if (inflight == DEPTH || (inflight > 0 && launched == 0)) {
                    if (nsb_engine_step_end(e) < 0) { fail("nsb_engine_step_end"); break; }
                    --inflight;
                    if (!pop_all()) break;
                }
                if (pushed == 0 && launched == 0 && inflight == 0) break;
            }
        }
        if (!ws.error.empty()) break;
        if (!pop_all()) break;
        nsb_stats st1;
        nsb_engine_get_stats(e, &st1);
        ws.chunks += st1.chunks - st0.chunks; ws.steps += st1.steps - st0.steps; ws.launches += st1.kernel_launches - st0.kernel_launches;
        ws.device_ms += st1.device_ms - st0.device_ms;
        for (int i = 0; i < n; ++i) {
            Result& r = results[mine[w0 + i]];
            r.chunks = nsb_stream_chunks(e, ids[i]);
            ws.audio_s += (double)r.chunks * shift / 16000.0;
            std::vector<char> text(r.tokens.size() * 9 + 16);
            const int nb = nsb_detokenize(e, r.tokens.data(), (int)r.tokens.size(), text.data(), (int)text.size());
            if (nb < 0) { fail("nsb_detokenize"); break; }
            r.text.assign(text.data(), (size_t)nb);
            nsb_stream_close(e, ids[i]);
        }
    }
    ws.wall_s = seconds_since(t_run);
    nsb_engine_destroy(e);
}

}  // namespace

int main(int argc, char** argv) {
    if (argc < 2) return usage(argv[0]);
    Options opt;
    std::vector<std::string> files;
    int synth_n = 0; double synth_s = 0;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&](const char* what) -> const char* {
            if (i + 1 >= argc) { fprintf(stderr, "Error: %s needs a value\n", what); exit(usage(argv[0])); }
            return argv[++i];
        };
        if (a == "--right-context") opt.right_context = atoi(next("--right-context"));
        else if (a == "--compute") { opt.compute = parse_compute(next("--compute")); if (opt.compute < 0) return usage(argv[0]); }
        else if (a == "--kv") { opt.kv = parse_kv(next("--kv")); if (opt.kv < 0) return usage(argv[0]); }
        else if (a == "--gpus") { const int g = atoi(next("--gpus")); if (g < 1) return usage(argv[0]); opt.devices.clear(); for (int d = 0; d < g; ++d) opt.devices.push_back(d); }
        else if (a == "--devices") {
            opt.devices.clear();
            std::string v = next("--devices");
            for (size_t p = 0; p <= v.size();) { const size_t q = std::min(v.find(',', p), v.size()); if (q > p) opt.devices.push_back(atoi(v.substr(p, q - p).c_str())); p = q + 1; }
            if (opt.devices.empty()) return usage(argv[0]);
        }
        else if (a == "--max-streams") { opt.max_streams = atoi(next("--max-streams")); if (opt.max_streams < 1) return usage(argv[0]); }
        else if (a == "--realtime") opt.realtime = true;
        else if (a == "--flush") opt.flush = true;
        else if (a == "--warmup") opt.warmup = true;
        else if (a == "--tokens") opt.print_tokens = true;
        else if (a == "--list") {
            const char* p = next("--list");
            std::ifstream f(p);
            if (!f) { fprintf(stderr, "Failed to open list file: %s\n", p); return 1; }
            for (std::string line; std::getline(f, line);) if (!line.empty()) files.push_back(line);
        }
        else if (a == "--synthetic") { synth_n = atoi(next("--synthetic")); synth_s = atof(next("--synthetic")); if (synth_n < 1 || synth_s <= 0) return usage(argv[0]); }
        else if (a == "--help" || a == "-h") { usage(argv[0]); return 0; }
        else if (a.size() > 2 && a[0] == '-' && a[1] == '-') { fprintf(stderr, "Error: unknown option %s\n", a.c_str()); return usage(argv[0]); }
        else if (opt.model.empty()) opt.model = a;
        else files.push_back(a);
    }
    if (opt.model.empty() || (files.empty() && synth_n == 0)) return usage(argv[0]);
    if (opt.right_context != 0 && opt.right_context != 1 && opt.right_context != 6 && opt.right_context != 13)
        fprintf(stderr, "Warning: right_context=%d is not standard. Valid values: 0, 1, 6, 13\n", opt.right_context);   // transcribe_stream.cpp:80-83

    nsb_model_info info;
    if (nsb_gguf_probe(opt.model.c_str(), &info) != NSB_OK) { fprintf(stderr, "Failed to load model: %s\n", nsb_last_error()); return 1; }

    // what NSB_COMPUTE_AUTO resolves to (engine.cu: F32 file -> fp32, F16 / Q4_0 -> fp16, Q8_0 -> fused dequant); the K/V ring follows
    // the arithmetic unless --kv says otherwise (fp32 ring in strict fp32, 16-bit ring in the 16-bit / Q8_0 modes)
    const int resolved = opt.compute != NSB_COMPUTE_AUTO ? opt.compute
                         : info.weight_type == 0 ? NSB_COMPUTE_F32 : info.weight_type == 8 ? NSB_COMPUTE_Q8_0 : NSB_COMPUTE_F16;
    if (opt.kv < 0) opt.kv = resolved == NSB_COMPUTE_F32 ? NSB_KV_F32 : resolved == NSB_COMPUTE_BF16 ? NSB_KV_BF16 : NSB_KV_F16;

    std::vector<Input> inputs;
    for (const std::string& p : files) {
        Input in; in.name = p;
        if (!read_pcm(p, in.pcm)) { fprintf(stderr, "Failed to open audio file: %s\n", p.c_str()); return 1; }
        inputs.push_back(std::move(in));
    }
    for (int s = 0; s < synth_n; ++s) inputs.push_back(Input{"synthetic:" + std::to_string(s), synthetic_pcm((uint32_t)s, synth_s)});

    const int G = (int)opt.devices.size();
    std::vector<std::vector<int>> share(G);
    for (size_t s = 0; s < inputs.size(); ++s) share[s % (size_t)G].push_back((int)s);       // stream s -> GPU s mod G, sticky
    std::vector<Result> results(inputs.size());
    std::vector<WorkerStats> stats(G);
    fprintf(stderr, "Model: %s (%d layers, weight type %d)\nStreams: %zu over %d GPU(s), right_context=%d (%d ms chunks)%s%s\n", opt.model.c_str(),
            info.n_layers, info.weight_type, inputs.size(), G, opt.right_context, 80 * (opt.right_context + 1), opt.realtime ? ", real-time pacing" : "",
            opt.flush ? ", tail flush" : "");

    const auto t0 = Clock::now();
    std::vector<std::thread> threads;
    for (int g = 0; g < G; ++g)
        threads.emplace_back(run_worker, std::cref(opt), opt.devices[g], std::cref(inputs), std::cref(share[g]), std::ref(results), std::ref(stats[g]));
    for (auto& t : threads) t.join();
    const double wall = seconds_since(t0);

    int rc = 0;
    for (const WorkerStats& ws : stats)
        if (!ws.error.empty()) { fprintf(stderr, "GPU %d: %s\n", ws.device, ws.error.c_str()); rc = ws.error.rfind("Failed to load model", 0) == 0 ? 1 : 2; }
    if (rc) return rc;

    for (size_t s = 0; s < inputs.size(); ++s) {          // host-side gather, input order
        printf("%zu\t%s\t%s", s, inputs[s].name.c_str(), results[s].text.c_str());
        if (opt.print_tokens) { printf("\t"); for (size_t k = 0; k < results[s].tokens.size(); ++k) printf(k ? " %d" : "%d", results[s].tokens[k]); }
        printf("\n");
    }
    fflush(stdout);

    double audio = 0, run_wall = 0; long long chunks = 0; std::vector<double> lat;
    for (const WorkerStats& ws : stats) {
        if (!ws.streams) continue;
        fprintf(stderr, "GPU %d: %d streams in %d wave(s), %lld stream-chunks in %lld steps, %.1f s audio, load %.2f s, run %.3f s (device %.1f ms), %.1f RTFx\n",
                ws.device, ws.streams, ws.waves, ws.chunks, ws.steps, ws.audio_s, ws.load_s, ws.wall_s, ws.device_ms, ws.wall_s > 0 ? ws.audio_s / ws.wall_s : 0.0);
        audio += ws.audio_s; chunks += ws.chunks; run_wall = std::max(run_wall, ws.wall_s);
        lat.insert(lat.end(), ws.latency_ms.begin(), ws.latency_ms.end());
    }
    fprintf(stderr, "Chunks processed:    %lld\nAudio duration:      %.2f sec\nProcessing time:     %.3f sec (%.3f sec with model load)\n", chunks, audio, run_wall, wall);
    if (run_wall > 0 && audio > 0) fprintf(stderr, "Real-time factor:    %.5fx (RTFx %.1f)\n", run_wall / audio, audio / run_wall);
    if (!lat.empty()) {
        std::sort(lat.begin(), lat.end());
        fprintf(stderr, "Chunk latency:       p50 %.3f ms, p99 %.3f ms (%zu steps)\n", lat[lat.size() / 2], lat[std::min(lat.size() - 1, (size_t)(0.99 * lat.size()))], lat.size());
    }
    return 0;
}
