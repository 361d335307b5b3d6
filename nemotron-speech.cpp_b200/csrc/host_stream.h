// host_stream.h -- host-side bookkeeping of one audio stream (pure C++, no CUDA): the buffered PCM, which chunk is due, and the PCM
// row a step stages for it. Header-only so that the CPU test suite can drive it against the CPU checker's chunk arithmetic
// (tests/test_host_api.py) -- the engine (engine.cu) uses exactly these functions.
//
// Geometry (reference: src/preprocessor.cpp:220-221,320-328; src/nemo-stream.h:65-100; src/nemo-stream.cpp:1094-1127), T = 1 + R:
//   the stream is left-padded with 256 zeros, mel frame t covers padded samples [160 t, 160 t + 512) and pre-emphasis needs one
//   sample of look-back (x[-1] = 0 at stream start); chunk c consumes the 8T NEW mel frames [8T c, 8T (c+1)) -- the 9 frames of
//   left context before them stay on the device -- so it is complete once raw sample 160 (8T (c+1) - 1) + 255 has arrived, and
//   its PCM row is raw[start - 1 .. start - 1 + row_len) with start = 1280 T c - 256 and row_len = 1280 T + 353.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <deque>
#include <vector>

namespace nsb {

constexpr int HS_HOP = 160, HS_NFFT = 512;

struct HostStream {
    bool open = false;
    std::vector<int16_t> buf;      // raw samples: buf[head] is absolute sample index `base`
    size_t head = 0;               // consumed prefix of buf (dropped lazily: a long one-shot push must not memmove per chunk)
    long long base = 0, n_pushed = 0, chunk_idx = 0, chunks_done = 0;   // chunk_idx: next chunk to LAUNCH; chunks_done: chunks whose tokens were collected
    std::deque<int32_t> tokens;
};

inline int hs_row_len(int T) { return 8 * T * HS_HOP + (HS_NFFT - HS_HOP) + 1; }          // 1280 T + 353 samples per stream-step

inline void hs_clear(HostStream& h) { std::vector<int16_t>().swap(h.buf); h.head = 0; h.base = 0; h.n_pushed = 0; h.chunk_idx = 0; h.chunks_done = 0; h.tokens.clear(); }

// stream closed: give the PCM buffer back (tokens already collected stay poppable)
inline void hs_release_pcm(HostStream& h) { std::vector<int16_t>().swap(h.buf); h.head = 0; h.base = h.n_pushed; }

inline void hs_push(HostStream& h, const int16_t* pcm, int n) {
    h.buf.insert(h.buf.end(), pcm, pcm + n);
    h.n_pushed += n;
}

// Chunk c needs mel frames up to 8T(c+1)-1, i.e. padded samples up to 160*(8T(c+1)-1)+512, i.e. raw samples
// n >= 160*(8T(c+1)-1) + 256  (preprocessor.cpp:320-328 frame count + nemo-stream.cpp:1094-1102 chunk gate)
inline bool hs_ready(const HostStream& h, int T) {
    return h.open && h.n_pushed >= (long long)HS_HOP * (8LL * T * (h.chunk_idx + 1) - 1) + HS_NFFT / 2;
}

// row = raw[start-1 .. start + 1280T + 352), start = 1280 T c - 256; negative indices are the 256-zero left pad / x[-1] = 0
inline void hs_stage_row(const HostStream& h, int T, int rl, int16_t* dst) {
    const long long start = 8LL * T * HS_HOP * h.chunk_idx - HS_NFFT / 2 - 1;
    const int zeros = start < 0 ? (int)std::min<long long>(-start, rl) : 0;
    if (zeros) memset(dst, 0, (size_t)zeros * sizeof(int16_t));
    if (zeros < rl) memcpy(dst + zeros, h.buf.data() + h.head + (size_t)(start + zeros - h.base), (size_t)(rl - zeros) * sizeof(int16_t));
}

// the chunk was launched: ready() now asks for the NEXT chunk; drop the samples no later chunk needs (the next row starts at
// 1280 T c' - 257)
inline void hs_launched(HostStream& h, int T) {
    h.chunk_idx += 1;
    const long long keep_from = std::max(0LL, 8LL * T * HS_HOP * h.chunk_idx - HS_NFFT / 2 - 1);
    if (keep_from > h.base) { h.head += (size_t)(keep_from - h.base); h.base = keep_from; }
    // compact only when the dead prefix outweighs the live samples: O(1) amortised per sample however the audio was pushed
    if (h.head >= 4096 && 2 * h.head >= h.buf.size()) { h.buf.erase(h.buf.begin(), h.buf.begin() + (std::ptrdiff_t)h.head); h.head = 0; }
}

}  // namespace nsb
