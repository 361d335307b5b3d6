// preprocessor.h -- drop-in for the reference's src/preprocessor.h (same function names and argument
// meaning), backed by the B200 log-mel kernel. The reference keeps ONE stateful CPU preprocessor per
// nemo_context; here the state (sample tail, pre-emphasis carry) lives with each stream inside the
// engine, so this object only carries what the stand-alone API needs.
#ifndef NEMO_PREPROCESSOR_H
#define NEMO_PREPROCESSOR_H

#include <cstddef>
#include <cstdint>
#include <vector>

struct nemo_preprocessor;

// src/preprocessor.h:20-23 -- file-path variant. The reference's version leaves the window at 400 taps
// and then indexes 512 (out of bounds, preprocessor.cpp:227-268 vs :296-299); it is rejected here.
struct nemo_preprocessor* nemo_preprocessor_init(const char* filterbank_path, const char* window_path);

// src/preprocessor.h:28-33 -- only validates sizes (128*257 filterbank, 400-tap window) and keeps a copy.
struct nemo_preprocessor* nemo_preprocessor_init_from_data(const float* filterbank_data, size_t filterbank_size,
                                                           const float* window_data, size_t window_size);
void nemo_preprocessor_free(struct nemo_preprocessor* pp);

// src/preprocessor.h:41-46 -- stateful PCM -> log-mel [n_frames, 128]. Needs a bound engine
// (nemo_preprocessor_bind, called by nemo_init*); returns 0 frames and warns otherwise.
size_t nemo_preprocessor_process(struct nemo_preprocessor* pp, const int16_t* audio, size_t n_samples, std::vector<float>& mel_out);

// src/preprocessor.h:49-52
size_t nemo_preprocessor_get_n_frames(struct nemo_preprocessor* pp, size_t n_samples);

#endif
