// kernels_decode.cu -- batched RNN-T greedy decode as ONE persistent cooperative kernel (sm_100a).
//
// Reference behaviour being replaced: decode_one_step + the per-frame loop of process_mel_chunk_streaming
// (src/nemo-stream.cpp:788-878, :1035-1047) on top of build_decoder_step / build_lstm_cell / build_joint
// (src/nemo-ggml.cpp:503-542, :1013-1100). The reference launches one graph and does 6-8 host<->device
// copies PER SYMBOL PER STREAM; here all streams of a step are decoded by one kernel launch with no host
// round-trips: the grid loops over "rounds" (one symbol evaluation for every still-active stream), with
// grid-wide barriers between the phases, and weights are read once per round for all streams.
//
// Semantics kept exactly (nemo-stream.cpp:813-875): up to 10 symbols per encoder frame; argmax = lowest index
// among maxima; blank => next frame, LSTM state untouched; non-blank => emit, prev_token = token, commit h', c'.
// The reference re-runs the LSTM for every evaluation and throws the result away on blank; since
// (prev_token, h, c) only change on emission the candidate (h', c', joint.pred projection) is cached per
// stream and recomputed only after an emission -- arithmetic-identical.
//
// Work split per round (grid = one CTA per SM):
//   LSTM   : each CTA owns a slice of the 640 hidden units (all 4 gates of a unit => the cell update is local)
//   pred   : each CTA owns a slice of the 640 joint.pred rows
//   joint  : each CTA owns a slice of the 1025 vocabulary rows and produces a partial (max, argmax) per stream
//   decide : one warp per stream reduces the partials and applies the blank / emit rule
#include <cooperative_groups.h>

#include "kernels.cuh"

namespace cg = cooperative_groups;

namespace nsb {

namespace {
constexpr int NT = 256, NW = NT / 32;
constexpr int BCH = 64;              // streams processed per shared-memory pass
constexpr int MAX_UNITS = 8;         // hidden units / rows per CTA (640 / 148 -> 5)
constexpr int EL = HID / 32;         // 20 elements of a 640-vector per lane

// weights: never written while the engine runs -> read-only (non-coherent) path is safe
__device__ __forceinline__ void load_vec_ro(const float* p, float (&r)[EL], int lane) {
#pragma unroll
    for (int e = 0; e < EL; ++e) r[e] = __ldg(p + lane + 32 * e);
}
// state written by other CTAs earlier in this kernel (cand_h, h, dec_proj): plain coherent loads only
__device__ __forceinline__ void load_vec(const float* p, float (&r)[EL], int lane) {
#pragma unroll
    for (int e = 0; e < EL; ++e) r[e] = p[lane + 32 * e];
}
__device__ __forceinline__ float dot_vec(const float (&w)[EL], const float (&x)[EL]) {
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < EL; ++e) s = fmaf(w[e], x[e], s);
    return s;
}

// one LSTM layer for all streams with need_lstm set; this CTA handles hidden units [u0, u1)
__device__ void lstm_layer_phase(const DecodeArgs& a, int layer, int u0, int u1, float* gates /*[BCH][4*MAX_UNITS]*/) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nu = u1 - u0;
    if (nu <= 0) return;
    const int nrows = 4 * nu;
    for (int bc = 0; bc < a.B; bc += BCH) {
        const int bn = min(BCH, a.B - bc);
        for (int r = warp; r < nrows; r += NW) {
            const int g = r / nu, u = u0 + r % nu, wrow = g * HID + u;
            float wi[EL], wh[EL];
            load_vec_ro(a.w.w_ih[layer] + (size_t)wrow * HID, wi, lane);
            load_vec_ro(a.w.w_hh[layer] + (size_t)wrow * HID, wh, lane);
            const float bias2 = a.w.b_ih[layer][wrow];
            const float bias3 = a.w.b_hh[layer][wrow];
            for (int bl = 0; bl < bn; ++bl) {
                const int b = bc + bl;
                if (!a.need_lstm[b]) continue;
                const int slot = a.slot_of_b[b];
                const float* x = layer == 0 ? a.w.embed + (size_t)a.s.prev_token[slot] * HID      // nemo-stream.cpp:825-828
                                            : a.s.cand_h + (size_t)slot * 2 * HID;               // layer-1 input = layer-0 h'
                const float* h = a.s.h + (size_t)slot * 2 * HID + layer * HID;
                float xv[EL], hv[EL];
                load_vec(x, xv, lane); load_vec(h, hv, lane);
                const float si = warp_sum(dot_vec(wi, xv)), sh = warp_sum(dot_vec(wh, hv));
                if (lane == 0) gates[bl * 4 * MAX_UNITS + g * MAX_UNITS + (u - u0)] = ((si + sh) + bias2) + bias3;   // nemo-ggml.cpp:518-522
            }
        }
        __syncthreads();
        for (int e = tid; e < bn * nu; e += NT) {                                 // cell update, gate order i,f,g,o :526-541
            const int bl = e / nu, uu = e % nu, b = bc + bl;
            if (!a.need_lstm[b]) continue;
            const int slot = a.slot_of_b[b];
            const float* gt = gates + bl * 4 * MAX_UNITS;
            const float ig = sigmoid_exact(gt[0 * MAX_UNITS + uu]), fg = sigmoid_exact(gt[1 * MAX_UNITS + uu]);
            const float gg = tanhf(gt[2 * MAX_UNITS + uu]), og = sigmoid_exact(gt[3 * MAX_UNITS + uu]);
            const size_t o = (size_t)slot * 2 * HID + layer * HID + u0 + uu;
            const float cn = fg * a.s.c[o] + ig * gg;
            a.s.cand_c[o] = cn;
            a.s.cand_h[o] = og * tanhf(cn);
        }
        __syncthreads();
    }
}

// joint.pred projection of the candidate decoder output (nemo-ggml.cpp:1086-1087); rows [j0, j1)
__device__ void pred_phase(const DecodeArgs& a, int j0, int j1) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int j = j0 + warp; j < j1; j += NW) {
        float w[EL];
        load_vec_ro(a.w.pred_w + (size_t)j * HID, w, lane);
        const float bias = a.w.pred_b[j];
        for (int b = 0; b < a.B; ++b) {
            if (!a.need_lstm[b]) continue;
            const int slot = a.slot_of_b[b];
            float x[EL];
            load_vec(a.s.cand_h + (size_t)slot * 2 * HID + HID, x, lane);
            const float s = warp_sum(dot_vec(w, x));
            if (lane == 0) a.s.dec_proj[(size_t)slot * HID + j] = s + bias;
        }
    }
}

__global__ void __launch_bounds__(NT, 1) rnnt_decode_kernel(const DecodeArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ float s_gates[BCH * 4 * MAX_UNITS];
    __shared__ float s_val[NW][BCH];
    __shared__ int s_idx[NW][BCH];
    const int nblk = gridDim.x, blk = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int B = a.B, T = a.T;
    const int upb = (HID + nblk - 1) / nblk;                                      // hidden units (and pred rows) per CTA
    const int u0 = min(HID, blk * upb), u1 = min(HID, u0 + upb);
    const int vpb = (VOCAB + nblk - 1) / nblk;                                    // vocabulary rows per CTA
    const int v0 = min(VOCAB, blk * vpb), v1 = min(VOCAB, v0 + vpb);

    for (int b = blk * NT + tid; b < B; b += nblk * NT) {
        a.frame_idx[b] = 0; a.sym_cnt[b] = 0; a.out_count[b] = 0;
        a.need_lstm[b] = a.s.cand_valid[a.slot_of_b[b]] ? 0 : 1;
    }
    if (blk == 0 && tid == 0 && a.logits_tap_n) *a.logits_tap_n = 0;
    grid.sync();

    for (;;) {
        // ---------------- prediction network for streams whose candidate is stale ----------------
        int need = 0;
        for (int b = tid; b < B; b += NT) need |= a.need_lstm[b];
        if (__syncthreads_or(need)) {
            lstm_layer_phase(a, 0, u0, u1, s_gates);
            grid.sync();
            lstm_layer_phase(a, 1, u0, u1, s_gates);
            grid.sync();
            pred_phase(a, u0, u1);
            grid.sync();
            if (blk == 0)
                for (int b = tid; b < B; b += NT)
                    if (a.need_lstm[b]) { a.need_lstm[b] = 0; a.s.cand_valid[a.slot_of_b[b]] = 1; }
        }
        // ---------------- joint + partial argmax over this CTA's vocabulary slice ----------------
        int act = 0;
        for (int b = tid; b < B; b += NT) act |= (a.frame_idx[b] < T);
        if (!__syncthreads_or(act)) break;                                        // uniform across the grid
        for (int bc = 0; bc < B; bc += BCH) {
            const int bn = min(BCH, B - bc);
            for (int e = tid; e < NW * BCH; e += NT) { (&s_val[0][0])[e] = -INFINITY; (&s_idx[0][0])[e] = 0x7fffffff; }
            __syncthreads();
            for (int v = v0 + warp; v < v1; v += NW) {
                float w[EL];
                load_vec_ro(a.w.out_w + (size_t)v * JOINT, w, lane);
                const float bias = a.w.out_b[v];
                for (int bl = 0; bl < bn; ++bl) {
                    const int b = bc + bl, f = a.frame_idx[b];
                    if (f >= T) continue;
                    const int slot = a.slot_of_b[b];
                    const float* ep = a.enc_proj + ((size_t)b * T + f) * JOINT;
                    const float* dp = a.s.dec_proj + (size_t)slot * JOINT;
                    float s = 0.f;
#pragma unroll
                    for (int e = 0; e < EL; ++e) s = fmaf(w[e], fmaxf(ep[lane + 32 * e] + dp[lane + 32 * e], 0.f), s);   // relu(enc+pred) :1092-1093
                    s = warp_sum(s) + bias;
                    if (lane == 0) {
                        if (s > s_val[warp][bl]) { s_val[warp][bl] = s; s_idx[warp][bl] = v; }   // v ascending within a warp
                        if (a.logits_tap && b == 0) { const int n = *a.logits_tap_n; if (n < a.logits_tap_cap) a.logits_tap[(size_t)n * VOCAB + v] = s; }
                    }
                }
            }
            __syncthreads();
            for (int bl = tid; bl < bn; bl += NT) {
                float bv = -INFINITY; int bi = 0x7fffffff;
#pragma unroll
                for (int w8 = 0; w8 < NW; ++w8) {
                    const float v = s_val[w8][bl]; const int i = s_idx[w8][bl];
                    if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
                }
                a.part_val[(size_t)(bc + bl) * nblk + blk] = bv;
                a.part_idx[(size_t)(bc + bl) * nblk + blk] = bi;
            }
            __syncthreads();
        }
        grid.sync();
        // ---------------- decision: one warp per stream ----------------
        for (int b = blk * NW + warp; b < B; b += nblk * NW) {
            if (a.frame_idx[b] >= T) continue;
            float bv = -INFINITY; int bi = 0x7fffffff;
            for (int p = lane; p < nblk; p += 32) {
                const float v = a.part_val[(size_t)b * nblk + p]; const int i = a.part_idx[(size_t)b * nblk + p];
                if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float v = __shfl_xor_sync(0xffffffffu, bv, o); const int i = __shfl_xor_sync(0xffffffffu, bi, o);
                if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
            }
            const int slot = a.slot_of_b[b];
            if (bi != BLANK) {                                                    // emit: commit candidate state (:869-874)
                for (int e = lane; e < 2 * HID; e += 32) {
                    a.s.h[(size_t)slot * 2 * HID + e] = a.s.cand_h[(size_t)slot * 2 * HID + e];
                    a.s.c[(size_t)slot * 2 * HID + e] = a.s.cand_c[(size_t)slot * 2 * HID + e];
                }
            }
            if (lane == 0) {
                if (b == 0 && a.logits_tap_n) *a.logits_tap_n += 1;
                if (bi == BLANK) { a.frame_idx[b] += 1; a.sym_cnt[b] = 0; }       // blank: next frame, state untouched (:856-859)
                else {
                    const int n = a.out_count[b];
                    a.out_tokens[(size_t)b * MAX_SYMBOLS * T + n] = bi; a.out_count[b] = n + 1;
                    a.s.prev_token[slot] = bi; a.s.cand_valid[slot] = 0; a.need_lstm[b] = 1;
                    const int sc = a.sym_cnt[b] + 1;
                    if (sc >= MAX_SYMBOLS) { a.frame_idx[b] += 1; a.sym_cnt[b] = 0; } else a.sym_cnt[b] = sc;   // :813
                }
            }
        }
        grid.sync();
    }
}
}  // namespace

static int g_decode_grid = 0;
static int decode_grid() {
    if (g_decode_grid == 0) {
        int dev = 0, sms = 0, per_sm = 0;
        NSB_CUDA(cudaGetDevice(&dev));
        NSB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        NSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rnnt_decode_kernel, NT, 0));
        if (per_sm < 1) throw CudaError("decode kernel cannot be resident");
        g_decode_grid = sms;                                                      // one CTA per SM: co-residency guaranteed
    }
    return g_decode_grid;
}
size_t decode_scratch_parts(int B) { return (size_t)B * decode_grid(); }

int launch_decode(const DecodeArgs& a, cudaStream_t st) {
    const int grid = decode_grid();
    if (HID > grid * MAX_UNITS) throw CudaError("decode: too few SMs for the unit partition");
    void* args[] = {(void*)&a};
    NSB_CUDA(cudaLaunchCooperativeKernel((void*)rnnt_decode_kernel, dim3(grid), dim3(NT), args, 0, st));
    return grid;
}

}  // namespace nsb
