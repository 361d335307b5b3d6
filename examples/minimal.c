/* examples/minimal.c -- the C ABI from plain C: one stream, push PCM, step, pop tokens, detokenise.
 *   gcc -std=c99 -Iinclude examples/minimal.c -Lnemotron-speech.cpp_b200 -lnsb200 -Wl,-rpath,$PWD/nemotron-speech.cpp_b200 -o minimal
 *   ./minimal model.gguf audio.pcm [right_context]
 */
#include <stdio.h>
#include <stdlib.h>

#include "nsb200.h"

int main(int argc, char** argv) {
    nsb_engine_config cfg;
    nsb_engine* e = NULL;
    nsb_model_info info;
    FILE* f;
    short* pcm;
    long n;
    int s, nt;
    int32_t tok[4096];
    char text[8 * 4096 + 1];
    if (argc < 3) { fprintf(stderr, "Usage: %s model.gguf audio.pcm [right_context]\n", argv[0]); return 1; }
    if (nsb_gguf_probe(argv[1], &info) != NSB_OK) { fprintf(stderr, "Failed to load model: %s\n", nsb_last_error()); return 1; }
    f = fopen(argv[2], "rb");
    if (!f) { fprintf(stderr, "Failed to open audio file: %s\n", argv[2]); return 1; }
    fseek(f, 0, SEEK_END); n = ftell(f) / 2; fseek(f, 0, SEEK_SET);
    pcm = (short*)malloc((size_t)(n > 0 ? n : 1) * 2);
    if (fread(pcm, 2, (size_t)n, f) != (size_t)n) { fprintf(stderr, "Failed to read audio data\n"); return 1; }
    fclose(f);
    nsb_default_config(&cfg);
    cfg.att_right_context = argc > 3 ? atoi(argv[3]) : 13;
    cfg.max_streams = 1;
    if (nsb_engine_create(argv[1], &cfg, &e) != NSB_OK) { fprintf(stderr, "Failed to create engine: %s\n", nsb_last_error()); return 1; }
    s = nsb_stream_open(e);
    nsb_stream_push_pcm(e, s, pcm, (int)n);
    while (nsb_engine_step(e) > 0) {}
    nt = nsb_stream_pop_tokens(e, s, tok, 4096);
    if (nsb_detokenize(e, tok, nt, text, (int)sizeof text) < 0) { fprintf(stderr, "%s\n", nsb_last_error()); return 1; }
    printf("%s\n", text);
    fprintf(stderr, "%d layers, %d chunks, %d tokens\n", info.n_layers, nsb_stream_chunks(e, s), nt);
    nsb_stream_close(e, s);
    nsb_engine_destroy(e);
    free(pcm);
    return 0;
}
