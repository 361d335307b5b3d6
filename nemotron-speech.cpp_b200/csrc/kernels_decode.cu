// kernels_decode.cu -- batched RNN-T greedy decode as ONE persistent cooperative kernel (sm_100a).
//
// Reference behaviour being replaced: decode_one_step + the per-frame loop of process_mel_chunk_streaming
// (src/nemo-stream.cpp:788-878, :1035-1047) on top of build_decoder_step / build_lstm_cell / build_joint
// (src/nemo-ggml.cpp:503-542, :1013-1100). The reference launches one graph and does 6-8 host<->device
// copies PER SYMBOL PER STREAM; here all streams of a step are decoded by one kernel launch with no host
// round-trips: the grid loops over "rounds" (one symbol evaluation for every still-active stream), with
// grid-wide barriers between the phases.
//
// Semantics kept exactly (nemo-stream.cpp:813-875): up to 10 symbols per encoder frame; argmax = lowest index
// among maxima; blank => next frame, LSTM state untouched; non-blank => emit, prev_token = token, commit h', c'.
// The reference re-runs the LSTM for every evaluation and throws the result away on blank; since
// (prev_token, h, c) only change on emission the candidate (h', c', joint.pred projection) is cached per
// stream and recomputed only after an emission -- arithmetic-identical.
//
// Work split per phase: the streams that need the phase are compacted into a list and cut into groups of
// GS = 16; the grid is arranged as (groups x row-slices). A CTA stages its group's input vectors in shared
// memory once, then each warp streams weight rows of its row-slice from L2 (coalesced, read-only path) and
// dots them against the 16 staged vectors. With few streams needing a phase (the common case after an
// emission) there is one group and all CTAs split the weight rows, so every weight byte is read once.
#include <cooperative_groups.h>

#include "kernels.cuh"

namespace cg = cooperative_groups;

namespace nsb {

namespace {
constexpr int NT = 256, NW = NT / 32;
constexpr int GS = 16;               // streams per group (staged in shared memory together)
constexpr int UB = 32;               // hidden units per gate block (4 * UB gate rows buffered)
constexpr int EL = HID / 32;         // 20 elements of a 640-vector per lane
constexpr int MAXB = 1024;           // max streams per step

struct DecSmem {
    float xs[GS][HID];               // staged input vectors (x or joint activations)
    float hs[GS][HID];               // staged recurrent vectors
    float gates[GS][4 * UB];
    float w_val[NW][GS]; int w_idx[NW][GS];
    int list[MAXB];                  // compacted stream list of the current phase
    int warp_cnt[NW];
    int g_slot[GS], g_aux[GS];       // per-group metadata hoisted out of the staging loops (slot; prev_token or frame row)
    int n_list;
};

__device__ __forceinline__ void load_row_ro(const float* p, float (&r)[EL], int lane) {      // weights: read-only path
#pragma unroll
    for (int e = 0; e < EL; ++e) r[e] = __ldg(p + lane + 32 * e);
}
__device__ __forceinline__ float dot_smem(const float (&w)[EL], const float* x, int lane) {
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < EL; ++e) s = fmaf(w[e], x[lane + 32 * e], s);
    return s;
}

// Sum 16 per-lane partials across the warp with 16 shuffles (instead of 16 x 5): every stage halves the values a lane
// keeps. On return lane l holds the complete sum of v[(l >> 1) & 15] in the return value.
__device__ __forceinline__ float reduce16(float (&v)[GS], int lane) {
#pragma unroll
    for (int half = 8, mask = 16; half >= 1; half >>= 1, mask >>= 1) {
        const bool hi = (lane & mask) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = hi ? v[i] : v[i + half];
            const float keep = hi ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
        }
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// block-wide stable compaction of {b : pred(b)} into sm.list; returns the count (block-uniform)
template <class Pred>
__device__ int compact(DecSmem& sm, int B, Pred pred) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int base = 0;
    for (int b0 = 0; b0 < B; b0 += NT) {
        const int b = b0 + tid;
        const bool f = b < B && pred(b);
        const unsigned m = __ballot_sync(0xffffffffu, f);
        if (lane == 0) sm.warp_cnt[warp] = __popc(m);
        __syncthreads();
        int off = base, tot = 0;
        for (int w = 0; w < NW; ++w) { if (w < warp) off += sm.warp_cnt[w]; tot += sm.warp_cnt[w]; }
        if (f) sm.list[off + __popc(m & ((1u << lane) - 1))] = b;
        base += tot;
        __syncthreads();
    }
    return base;
}

// grid arrangement for n listed streams: groups of GS x row slices
struct Part { int group, n_groups, rs, n_rs; };
__device__ __forceinline__ Part make_part(int n, int nblk, int blk) {
    Part p; p.n_groups = (n + GS - 1) / GS; p.n_rs = max(1, nblk / max(1, p.n_groups));
    p.group = blk / p.n_rs; p.rs = blk % p.n_rs; return p;
}

// one LSTM layer for the listed streams (gate order i,f,g,o: nemo-ggml.cpp:518-541)
__device__ void lstm_phase(const DecodeArgs& a, DecSmem& sm, int layer, int n, int nblk, int blk) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const Part p = make_part(n, nblk, blk);
    if (p.group >= p.n_groups) return;
    const int g0 = p.group * GS, gn = min(GS, n - g0);
    const int upb = (HID + p.n_rs - 1) / p.n_rs, u0 = min(HID, p.rs * upb), u1 = min(HID, u0 + upb);
    if (u0 >= u1) return;
    if (tid < gn) { const int slot = a.slot_of_b[sm.list[g0 + tid]]; sm.g_slot[tid] = slot; sm.g_aux[tid] = a.s.prev_token[slot]; }
    __syncthreads();
    for (int g = 0; g < gn; ++g) {                                                // stage x and h (independent coalesced loads)
        const int slot = sm.g_slot[g];
        const float* xsrc = layer == 0 ? a.w.embed + (size_t)sm.g_aux[g] * HID                   // nemo-stream.cpp:825-828
                                       : a.s.cand_h + (size_t)slot * 2 * HID;                    // layer-1 input = layer-0 h'
        const float* hsrc = a.s.h + (size_t)slot * 2 * HID + layer * HID;
        for (int k = tid; k < HID; k += NT) { sm.xs[g][k] = xsrc[k]; sm.hs[g][k] = hsrc[k]; }
    }
    __syncthreads();
    for (int ub = u0; ub < u1; ub += UB) {
        const int nu = min(UB, u1 - ub), nrows = 4 * nu;
        for (int r = warp; r < nrows; r += NW) {
            const int gate = r / nu, u = ub + r % nu, wrow = gate * HID + u;
            float wi[EL], wh[EL];
            load_row_ro(a.w.w_ih[layer] + (size_t)wrow * HID, wi, lane);
            load_row_ro(a.w.w_hh[layer] + (size_t)wrow * HID, wh, lane);
            const float b2 = __ldg(a.w.b_ih[layer] + wrow), b3 = __ldg(a.w.b_hh[layer] + wrow);
            float acc[GS];
#pragma unroll
            for (int g = 0; g < GS; ++g) acc[g] = g < gn ? dot_smem(wi, sm.xs[g], lane) + dot_smem(wh, sm.hs[g], lane) : 0.f;
            const float tot = reduce16(acc, lane);
            const int g = (lane >> 1) & 15;
            if (!(lane & 1) && g < gn) sm.gates[g][gate * UB + (u - ub)] = (tot + b2) + b3;
        }
        __syncthreads();
        for (int e = tid; e < gn * nu; e += NT) {
            const int g = e / nu, uu = e % nu, slot = sm.g_slot[g];
            const float ig = sigmoid_exact(sm.gates[g][0 * UB + uu]), fg = sigmoid_exact(sm.gates[g][1 * UB + uu]);
            const float gg = tanhf(sm.gates[g][2 * UB + uu]), og = sigmoid_exact(sm.gates[g][3 * UB + uu]);
            const size_t o = (size_t)slot * 2 * HID + layer * HID + ub + uu;
            const float cn = fg * a.s.c[o] + ig * gg;
            a.s.cand_c[o] = cn; a.s.cand_h[o] = og * tanhf(cn);
        }
        __syncthreads();
    }
}

// joint.pred projection of the candidate decoder output (nemo-ggml.cpp:1086-1087)
__device__ void pred_phase(const DecodeArgs& a, DecSmem& sm, int n, int nblk, int blk) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const Part p = make_part(n, nblk, blk);
    if (p.group >= p.n_groups) return;
    const int g0 = p.group * GS, gn = min(GS, n - g0);
    const int rpb = (JOINT + p.n_rs - 1) / p.n_rs, j0 = min(JOINT, p.rs * rpb), j1 = min(JOINT, j0 + rpb);
    if (j0 >= j1) return;
    if (tid < gn) sm.g_slot[tid] = a.slot_of_b[sm.list[g0 + tid]];
    __syncthreads();
    for (int g = 0; g < gn; ++g) {
        const float* src = a.s.cand_h + (size_t)sm.g_slot[g] * 2 * HID + HID;
        for (int k = tid; k < HID; k += NT) sm.xs[g][k] = src[k];
    }
    __syncthreads();
    for (int j = j0 + warp; j < j1; j += NW) {
        float w[EL]; load_row_ro(a.w.pred_w + (size_t)j * HID, w, lane);
        const float bias = __ldg(a.w.pred_b + j);
        float acc[GS];
#pragma unroll
        for (int g = 0; g < GS; ++g) acc[g] = g < gn ? dot_smem(w, sm.xs[g], lane) : 0.f;
        const float tot = reduce16(acc, lane);
        const int g = (lane >> 1) & 15;
        if (!(lane & 1) && g < gn) a.s.dec_proj[(size_t)sm.g_slot[g] * JOINT + j] = tot + bias;
    }
}

// joint network + partial argmax for the listed (active) streams (nemo-ggml.cpp:1092-1097)
__device__ void joint_phase(const DecodeArgs& a, DecSmem& sm, int n, int nblk, int blk) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const Part p = make_part(n, nblk, blk);
    if (p.group >= p.n_groups) return;
    const int g0 = p.group * GS, gn = min(GS, n - g0);
    const int vpb = (VOCAB + p.n_rs - 1) / p.n_rs, v0 = min(VOCAB, p.rs * vpb), v1 = min(VOCAB, v0 + vpb);
    if (tid < gn) { const int b = sm.list[g0 + tid]; sm.g_slot[tid] = a.slot_of_b[b]; sm.g_aux[tid] = b * a.T + a.frame_idx[b]; }
    __syncthreads();
    for (int g = 0; g < gn; ++g) {                                                // z = relu(enc_proj + pred_proj)
        const float* ep = a.enc_proj + (size_t)sm.g_aux[g] * JOINT; const float* dp = a.s.dec_proj + (size_t)sm.g_slot[g] * JOINT;
        for (int k = tid; k < JOINT; k += NT) sm.xs[g][k] = fmaxf(ep[k] + dp[k], 0.f);
    }
    float bv = -INFINITY; int bi = 0x7fffffff;                                 // lane keeps the running best of stream g = (lane >> 1) & 15
    const int gl = (lane >> 1) & 15;
    const bool tap = a.logits_tap && !(lane & 1) && gl < gn && sm.list[g0 + gl] == 0;
    __syncthreads();
    for (int v = v0 + warp; v < v1; v += NW) {                                    // v ascending per warp => first max wins
        float w[EL]; load_row_ro(a.w.out_w + (size_t)v * JOINT, w, lane);
        const float bias = __ldg(a.w.out_b + v);
        float acc[GS];
#pragma unroll
        for (int g = 0; g < GS; ++g) acc[g] = g < gn ? dot_smem(w, sm.xs[g], lane) : 0.f;
        const float s = reduce16(acc, lane) + bias;
        if (s > bv) { bv = s; bi = v; }
        if (tap) { const int ne = *a.logits_tap_n; if (ne < a.logits_tap_cap) a.logits_tap[(size_t)ne * VOCAB + v] = s; }
    }
    if (!(lane & 1)) { sm.w_val[warp][gl] = bv; sm.w_idx[warp][gl] = bi; }
    __syncthreads();
    if (tid < gn) {
        float v = -INFINITY; int i = 0x7fffffff;
#pragma unroll
        for (int w8 = 0; w8 < NW; ++w8) {
            const float vv = sm.w_val[w8][tid]; const int ii = sm.w_idx[w8][tid];
            if (vv > v || (vv == v && ii < i)) { v = vv; i = ii; }
        }
        const int b = sm.list[g0 + tid];
        a.part_val[(size_t)b * nblk + p.rs] = v; a.part_idx[(size_t)b * nblk + p.rs] = i;
    }
}

__global__ void __launch_bounds__(NT, 1) rnnt_decode_kernel(const DecodeArgs a) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ uint8_t dec_smem_raw[];
    DecSmem& sm = *reinterpret_cast<DecSmem*>(dec_smem_raw);
    const int nblk = gridDim.x, blk = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int B = a.B, T = a.T;

    for (int b = blk * NT + tid; b < B; b += nblk * NT) {
        a.frame_idx[b] = 0; a.sym_cnt[b] = 0; a.out_count[b] = 0;
        a.need_lstm[b] = a.s.cand_valid[a.slot_of_b[b]] ? 0 : 1;
    }
    if (blk == 0 && tid == 0 && a.logits_tap_n) *a.logits_tap_n = 0;
    grid.sync();

    for (;;) {
        // ---------------- prediction network for streams whose candidate is stale ----------------
        const int n_need = compact(sm, B, [&](int b) { return a.need_lstm[b] != 0; });
        if (n_need > 0) {
            lstm_phase(a, sm, 0, n_need, nblk, blk);
            grid.sync();
            lstm_phase(a, sm, 1, n_need, nblk, blk);
            grid.sync();
            pred_phase(a, sm, n_need, nblk, blk);
            grid.sync();
            if (blk == 0)
                for (int i = tid; i < n_need; i += NT) { const int b = sm.list[i]; a.need_lstm[b] = 0; a.s.cand_valid[a.slot_of_b[b]] = 1; }
            __syncthreads();
        }
        // ---------------- joint + partial argmax ----------------
        const int n_act = compact(sm, B, [&](int b) { return a.frame_idx[b] < T; });
        if (n_act == 0) break;                                                    // uniform across the grid
        joint_phase(a, sm, n_act, nblk, blk);
        const int n_rs = make_part(n_act, nblk, blk).n_rs;
        grid.sync();
        // ---------------- decision: one warp per active stream ----------------
        for (int li = blk * NW + warp; li < n_act; li += nblk * NW) {
            const int b = sm.list[li];
            float bv = -INFINITY; int bi = 0x7fffffff;
            for (int q = lane; q < n_rs; q += 32) {
                const float v = a.part_val[(size_t)b * nblk + q]; const int i = a.part_idx[(size_t)b * nblk + q];
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
        NSB_CUDA(cudaFuncSetAttribute(rnnt_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DecSmem)));
        NSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rnnt_decode_kernel, NT, sizeof(DecSmem)));
        if (per_sm < 1) throw CudaError("decode kernel cannot be resident");
        g_decode_grid = sms;                                                      // one CTA per SM: co-residency guaranteed
    }
    return g_decode_grid;
}
size_t decode_scratch_parts(int B) { return (size_t)B * decode_grid(); }

int launch_decode(const DecodeArgs& a, cudaStream_t st) {
    const int grid = decode_grid();
    if (a.B > MAXB) throw CudaError("decode: more than 1024 streams in one step");
    void* args[] = {(void*)&a};
    NSB_CUDA(cudaLaunchCooperativeKernel((void*)rnnt_decode_kernel, dim3(grid), dim3(NT), args, sizeof(DecSmem), st));
    return grid;
}

}  // namespace nsb
