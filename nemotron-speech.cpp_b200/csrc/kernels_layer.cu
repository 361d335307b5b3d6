// kernels_layer.cu -- non-GEMM kernels of the cache-aware FastConformer layer (sm_100a).
//
// Reference behaviour being replaced (src/nemo-stream.cpp):
//   build_layer_norm :552-563 | build_cached_rel_pos_mha + build_cached_rel_shift :391-545 |
//   build_cached_causal_conv1d :308-384 + GLU / LayerNorm / SiLU of the conv module :629-646 |
//   cache roll-over :222-238, :477-484 and cache_valid_len update :1018.
// Design differences (B200-first): the K/V cache is a ring of L+T rows per (stream, layer) that is
// appended in place (the reference rewrites 2x70x1024 floats per layer per chunk); the relative-position
// term gather-indexes a pre-projected table instead of the pad/reshape shift; invalid cache slots are
// skipped by count instead of a -1e9 additive mask; GLU + depthwise conv + LayerNorm + SiLU + cache
// update are one kernel.
#include "kernels.cuh"

namespace nsb {

// ------------------------------------------------------------------------------------------
// block-wide sum over 256 threads (8 warps)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum_256(float v, float* red /*[8]*/) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();                       // protect red[] reuse
    if (l == 0) red[w] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i];
    return t;
}

// LayerNorm over rows of 1024, eps 1e-5, two-pass (mean, then variance of centred values) like ggml_norm.
// One CTA (256 threads x 4 channels) per row.
__device__ __forceinline__ float4 load_x_reduced(float* x, size_t o, const PartialSum& ps, int rows) {
    float4 v = *(const float4*)(x + o);
    if (ps.n > 0) {                                                              // fold the split-K partials of the previous GEMM
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = 0; s < ps.n; ++s) {
            const float4 t = *(const float4*)(ps.part + (size_t)s * rows * D_MODEL + o);
            acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        }
        v.x += ps.alpha * acc.x; v.y += ps.alpha * acc.y; v.z += ps.alpha * acc.z; v.w += ps.alpha * acc.w;
        *(float4*)(x + o) = v;
    }
    return v;
}

__global__ void __launch_bounds__(256) layernorm_kernel(float* x, const float* __restrict__ g,
                                                        const float* __restrict__ b, void* y, int out_type, const PartialSum ps) {
    __shared__ float red[8];
    const int row = blockIdx.x, c = threadIdx.x * 4;
    const float4 v = load_x_reduced(x, (size_t)row * D_MODEL + c, ps, gridDim.x);
    const float mean = block_sum_256(v.x + v.y + v.z + v.w, red) * (1.0f / D_MODEL);
    const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
    const float var = block_sum_256(d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3, red) * (1.0f / D_MODEL);
    const float rs = 1.0f / sqrtf(var + 1e-5f);
    const float4 gg = *(const float4*)(g + c), bb = *(const float4*)(b + c);
    const size_t o = (size_t)row * D_MODEL + c;
    const float o0 = d0 * rs * gg.x + bb.x, o1 = d1 * rs * gg.y + bb.y, o2 = d2 * rs * gg.z + bb.z, o3 = d3 * rs * gg.w + bb.w;
    if (out_type == OUT_F32) *(float4*)((float*)y + o) = make_float4(o0, o1, o2, o3);
    else if (out_type == OUT_F16) { __half2* p = (__half2*)((__half*)y + o); p[0] = __floats2half2_rn(o0, o1); p[1] = __floats2half2_rn(o2, o3); }
    else { __nv_bfloat162* p = (__nv_bfloat162*)((__nv_bfloat16*)y + o); p[0] = __floats2bfloat162_rn(o0, o1); p[1] = __floats2bfloat162_rn(o2, o3); }
}
void launch_layernorm(float* x, int rows, const float* g, const float* b, void* y, int out_type, const PartialSum& ps, cudaStream_t st) {
    if (rows > 0) layernorm_kernel<<<rows, 256, 0, st>>>(x, g, b, y, out_type, ps);
}

// norm_out of layer l fused with norm_feed_forward1 of layer l+1: x <- LN1(x) (f32, in place); y2 <- LN2(x)
__global__ void __launch_bounds__(256) layernorm2_kernel(float* x, const float* __restrict__ g1, const float* __restrict__ b1,
                                                         const float* __restrict__ g2, const float* __restrict__ b2,
                                                         void* y2, int out_type, const PartialSum ps) {
    __shared__ float red[8];
    const int row = blockIdx.x, c = threadIdx.x * 4;
    const size_t o = (size_t)row * D_MODEL + c;
    float4 v = load_x_reduced(x, o, ps, gridDim.x);
    float mean = block_sum_256(v.x + v.y + v.z + v.w, red) * (1.0f / D_MODEL);
    float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
    float var = block_sum_256(d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3, red) * (1.0f / D_MODEL);
    float rs = 1.0f / sqrtf(var + 1e-5f);
    float4 gg = *(const float4*)(g1 + c), bb = *(const float4*)(b1 + c);
    v = make_float4(d0 * rs * gg.x + bb.x, d1 * rs * gg.y + bb.y, d2 * rs * gg.z + bb.z, d3 * rs * gg.w + bb.w);
    *(float4*)(x + o) = v;
    if (y2 == nullptr) return;                                                   // last layer: no following norm (uniform branch)
    mean = block_sum_256(v.x + v.y + v.z + v.w, red) * (1.0f / D_MODEL);
    d0 = v.x - mean; d1 = v.y - mean; d2 = v.z - mean; d3 = v.w - mean;
    var = block_sum_256(d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3, red) * (1.0f / D_MODEL);
    rs = 1.0f / sqrtf(var + 1e-5f);
    gg = *(const float4*)(g2 + c); bb = *(const float4*)(b2 + c);
    const float o0 = d0 * rs * gg.x + bb.x, o1 = d1 * rs * gg.y + bb.y, o2 = d2 * rs * gg.z + bb.z, o3 = d3 * rs * gg.w + bb.w;
    if (out_type == OUT_F32) *(float4*)((float*)y2 + o) = make_float4(o0, o1, o2, o3);
    else if (out_type == OUT_F16) { __half2* p = (__half2*)((__half*)y2 + o); p[0] = __floats2half2_rn(o0, o1); p[1] = __floats2half2_rn(o2, o3); }
    else { __nv_bfloat162* p = (__nv_bfloat162*)((__nv_bfloat16*)y2 + o); p[0] = __floats2bfloat162_rn(o0, o1); p[1] = __floats2bfloat162_rn(o2, o3); }
}
void launch_layernorm2(float* x, int rows, const float* g1, const float* b1, const float* g2, const float* b2, void* y2, int out_type,
                       const PartialSum& ps, cudaStream_t st) {
    if (rows > 0) layernorm2_kernel<<<rows, 256, 0, st>>>(x, g1, b1, g2, b2, y2, out_type, ps);
}

// ------------------------------------------------------------------------------------------
// Cached relative-position attention. One CTA per (head, batch row); 256 threads.
//   keys j = 0..K-1 (K = L+T): j < L are ring rows (oldest first), j >= L are this chunk's new rows
//   score[i][j] = ((q_i + u) . k_j + (q_i + v) . P[L + i - j]) / sqrt(128),  j >= L - valid_len
//   ctx[i] = softmax_j(score[i]) . v_j
// All global traffic happens in one bulk staging phase (K, V head slices of the ring, the head slice of the projected
// positional table, q) with 16-byte loads and no dependent chains; scores / softmax / context then run out of shared
// memory. Row strides are padded by 16 bytes so that 16-byte row-wise reads by consecutive threads are conflict-free.
// The CTA also appends this chunk's K/V rows for its head to the ring (positions (w + i) mod (L+T)).
// ------------------------------------------------------------------------------------------
template <int KV> struct KvT { using type = float; };
template <> struct KvT<1> { using type = __half; };
template <> struct KvT<2> { using type = __nv_bfloat16; };

template <typename E> struct AttnLayout {
    static constexpr int EPV = 16 / sizeof(E);                 // elements per 16-byte vector
    static constexpr int KSTR = D_HEAD + EPV;                  // padded row stride (elements) of K / V tiles
    static constexpr int PSTR = D_HEAD + 4;                    // padded row stride (floats) of the P tile and q rows
};

__device__ __forceinline__ void unpack16(const uint4& v, float (&f)[4], const float*) {
    f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y); f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
}
__device__ __forceinline__ void unpack16(const uint4& v, float (&f)[8], const __half*) {
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 t = __half22float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ void unpack16(const uint4& v, float (&f)[8], const __nv_bfloat16*) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}

template <int KV>
__global__ void __launch_bounds__(256) attention_kernel(const AttnArgs a) {
    using E = typename KvT<KV>::type;
    using Lay = AttnLayout<E>;
    constexpr int EPV = Lay::EPV, KSTR = Lay::KSTR, PSTR = Lay::PSTR;
    extern __shared__ __align__(16) uint8_t sm_raw[];
    const int T = a.T, K = ATT_L + T, Cap = K, n_rel = ATT_L + 2 * T - 1;
    E* Ks = reinterpret_cast<E*>(sm_raw);                       // [K][KSTR]
    E* Vs = Ks + (size_t)K * KSTR;                              // [K][KSTR]
    float* Ps = reinterpret_cast<float*>(Vs + (size_t)K * KSTR);   // [n_rel][PSTR]
    float* qu = Ps + (size_t)n_rel * PSTR;                      // [T][PSTR]
    float* qv = qu + (size_t)T * PSTR;                          // [T][PSTR]
    float* sc = qv + (size_t)T * PSTR;                          // [T][K]
    const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int slot = a.slot_of_b[b], w = a.ring_pos[slot], valid = a.valid_len[slot];
    const int first = ATT_L - valid;                                           // keys j < first are not yet valid (:982-992)
    const float* qkv = a.qkv + (size_t)b * T * 3 * D_MODEL;
    const size_t ring_base = (size_t)slot * a.slot_stride + h * D_HEAD;
    E* kring = reinterpret_cast<E*>(a.k_ring); E* vring = reinterpret_cast<E*>(a.v_ring);

    // ---- bulk staging: cached K/V rows (16-byte loads), positional slice, q; new rows go to smem AND to the ring ----
    constexpr int VPR = D_HEAD / EPV;                                          // 16-byte vectors per row
    for (int e = tid; e < (ATT_L - first) * VPR; e += 256) {
        const int j = first + e / VPR, c = (e % VPR) * EPV;
        const int rrow = (w + Cap - ATT_L + j) % Cap;
        const size_t g = ring_base + (size_t)rrow * D_MODEL + c;
        *reinterpret_cast<uint4*>(Ks + (size_t)j * KSTR + c) = *reinterpret_cast<const uint4*>(kring + g);
        *reinterpret_cast<uint4*>(Vs + (size_t)j * KSTR + c) = *reinterpret_cast<const uint4*>(vring + g);
    }
    for (int e = tid; e < n_rel * (D_HEAD / 4); e += 256) {
        const int r = e / (D_HEAD / 4), c = (e % (D_HEAD / 4)) * 4;
        *reinterpret_cast<float4*>(Ps + (size_t)r * PSTR + c) = *reinterpret_cast<const float4*>(a.pos_proj + (size_t)r * D_MODEL + h * D_HEAD + c);
    }
    for (int e = tid; e < T * D_HEAD; e += 256) {
        const int i = e / D_HEAD, d = e % D_HEAD;
        const float* row = qkv + (size_t)i * 3 * D_MODEL + h * D_HEAD + d;
        const float q = row[0];
        qu[i * PSTR + d] = q + a.bias_u[h * D_HEAD + d];                       // :503-507
        qv[i * PSTR + d] = q + a.bias_v[h * D_HEAD + d];
        const E kn = from_f32<E>(row[D_MODEL]), vn = from_f32<E>(row[2 * D_MODEL]);
        Ks[(size_t)(ATT_L + i) * KSTR + d] = kn; Vs[(size_t)(ATT_L + i) * KSTR + d] = vn;
        const size_t r = ring_base + (size_t)((w + i) % Cap) * D_MODEL + d;    // append (replaces concat + roll :465-484)
        kring[r] = kn; vring[r] = vn;
    }
    __syncthreads();

    // ---- scores: one thread per (query i, key j) pair, full 128-dim dot products out of shared memory ----
    const float scale = 0.08838834764831845f;                                   // 1/sqrt(128) :517
    const int nkeys = K - first;
    for (int p = tid; p < T * nkeys; p += 256) {
        const int i = p / nkeys, j = first + p % nkeys;
        const E* kr = Ks + (size_t)j * KSTR;
        const float* pr = Ps + (size_t)((ATT_L + i - j) + (T - 1)) * PSTR;      // rel = L + i - j
        const float* qur = qu + i * PSTR; const float* qvr = qv + i * PSTR;
        float s = 0.f;
#pragma unroll 4
        for (int c = 0; c < D_HEAD; c += EPV) {
            float kf[EPV];
            unpack16(*reinterpret_cast<const uint4*>(kr + c), kf, (const E*)nullptr);
#pragma unroll
            for (int u = 0; u < EPV; u += 4) {
                const float4 q4 = *reinterpret_cast<const float4*>(qur + c + u), v4 = *reinterpret_cast<const float4*>(qvr + c + u);
                const float4 p4 = *reinterpret_cast<const float4*>(pr + c + u);
                s = fmaf(q4.x, kf[u], s); s = fmaf(q4.y, kf[u + 1], s); s = fmaf(q4.z, kf[u + 2], s); s = fmaf(q4.w, kf[u + 3], s);
                s = fmaf(v4.x, p4.x, s); s = fmaf(v4.y, p4.y, s); s = fmaf(v4.z, p4.z, s); s = fmaf(v4.w, p4.w, s);
            }
        }
        sc[i * K + j] = s * scale;
    }
    __syncthreads();
    for (int i = warp; i < T; i += 8) {                                         // softmax over valid keys
        float mx = -INFINITY;
        for (int j = first + lane; j < K; j += 32) mx = fmaxf(mx, sc[i * K + j]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int j = first + lane; j < K; j += 32) { const float e = expf(sc[i * K + j] - mx); sc[i * K + j] = e; sum += e; }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
        for (int j = first + lane; j < K; j += 32) sc[i * K + j] *= inv;
    }
    __syncthreads();
    // ---- ctx[i][d] = sum_j p[i][j] v_j[d]; thread = (i parity, d) ----
    const int d = tid & 127;
    for (int i = tid >> 7; i < T; i += 2) {
        float acc = 0.f;
#pragma unroll 4
        for (int j = first; j < K; ++j) acc = fmaf(sc[i * K + j], to_f32(Vs[(size_t)j * KSTR + d]), acc);
        store_out(a.ctx, ((size_t)b * T + i) * D_MODEL + h * D_HEAD + d, acc, a.out_type);
    }
}

template <int KV>
static void launch_attention_t(const AttnArgs& a, cudaStream_t st) {
    using E = typename KvT<KV>::type;
    using Lay = AttnLayout<E>;
    const int K = ATT_L + a.T, n_rel = ATT_L + 2 * a.T - 1;
    const size_t smem = (size_t)2 * K * Lay::KSTR * sizeof(E) + (size_t)(n_rel + 2 * a.T) * Lay::PSTR * 4 + (size_t)a.T * K * 4;
    static size_t configured = 0;
    if (smem > configured) {
        NSB_CUDA(cudaFuncSetAttribute(attention_kernel<KV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    attention_kernel<KV><<<dim3(N_HEADS, a.B), 256, smem, st>>>(a);
}
void launch_attention(const AttnArgs& a, cudaStream_t st) {
    if (a.kv_dtype == 0) launch_attention_t<0>(a, st);
    else if (a.kv_dtype == 1) launch_attention_t<1>(a, st);
    else launch_attention_t<2>(a, st);
}

// ------------------------------------------------------------------------------------------
// Conv module core: GLU -> cached causal depthwise conv (k = 9) -> LayerNorm -> SiLU, cache updated in place.
// One CTA per batch row (stream), 256 threads x 4 channels, a 9-deep register window slides over time.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) conv_module_kernel(const ConvModArgs a) {
    __shared__ float red[8];
    const int b = blockIdx.x, c0 = threadIdx.x * 4, T = a.T;
    const int slot = a.slot_of_b[b];
    float* cache = a.conv_cache + (size_t)slot * a.slot_stride;
    float win[4][CONV_K];
    float wk[4][CONV_K];
#pragma unroll
    for (int k = 0; k < CONV_K; ++k) {
        const float4 w4 = *(const float4*)(a.dw_w + (size_t)k * D_MODEL + c0);
        wk[0][k] = w4.x; wk[1][k] = w4.y; wk[2][k] = w4.z; wk[3][k] = w4.w;
    }
#pragma unroll
    for (int k = 0; k < CONV_K - 1; ++k) {                                       // xp = [cache(8) || glu(T)] :323-328
        const float4 v = *(const float4*)(cache + (size_t)k * D_MODEL + c0);
        win[0][k] = v.x; win[1][k] = v.y; win[2][k] = v.z; win[3][k] = v.w;
    }
    const float4 g4 = *(const float4*)(a.ln_g + c0), b4 = *(const float4*)(a.ln_b + c0);
    const float lg[4] = {g4.x, g4.y, g4.z, g4.w}, lb[4] = {b4.x, b4.y, b4.z, b4.w};
    for (int t = 0; t < T; ++t) {
        const float* row = a.pw1 + ((size_t)b * T + t) * 2 * D_MODEL;
        const float4 av = *(const float4*)(row + c0), gv = *(const float4*)(row + D_MODEL + c0);
        win[0][8] = av.x * sigmoid_exact(gv.x); win[1][8] = av.y * sigmoid_exact(gv.y);   // GLU :629-636
        win[2][8] = av.z * sigmoid_exact(gv.z); win[3][8] = av.w * sigmoid_exact(gv.w);
        float cv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {                                            // :341-360
            float acc = win[u][0] * wk[u][0];
#pragma unroll
            for (int k = 1; k < CONV_K; ++k) acc = fmaf(win[u][k], wk[u][k], acc);
            cv[u] = acc;
        }
        const float mean = block_sum_256(cv[0] + cv[1] + cv[2] + cv[3], red) * (1.0f / D_MODEL);     // LN :643-645
        float dd[4]; float sq = 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) { dd[u] = cv[u] - mean; sq += dd[u] * dd[u]; }
        const float var = block_sum_256(sq, red) * (1.0f / D_MODEL);
        const float rs = 1.0f / sqrtf(var + 1e-5f);
        const size_t o = ((size_t)b * T + t) * D_MODEL + c0;
#pragma unroll
        for (int u = 0; u < 4; ++u) store_out(a.out, o + u, silu_exact(dd[u] * rs * lg[u] + lb[u]), a.out_type);   // SiLU :646
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int k = 0; k < CONV_K - 1; ++k) win[u][k] = win[u][k + 1];
    }
#pragma unroll
    for (int k = 0; k < CONV_K - 1; ++k)                                         // new cache = last 8 rows of xp :368-381
        *(float4*)(cache + (size_t)k * D_MODEL + c0) = make_float4(win[0][k], win[1][k], win[2][k], win[3][k]);
}
void launch_conv_module(const ConvModArgs& a, cudaStream_t st) {
    if (a.B > 0) conv_module_kernel<<<a.B, 256, 0, st>>>(a);
}

__global__ void advance_streams_kernel(const int* __restrict__ slot_of_b, int B, int T, int* ring_pos, int* valid_len) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int s = slot_of_b[b];
    ring_pos[s] = (ring_pos[s] + T) % (ATT_L + T);
    valid_len[s] = min(valid_len[s] + T, ATT_L);                                 // :1018
}
void launch_advance_streams(const int* slot_of_b, int B, int T, int* ring_pos, int* valid_len, cudaStream_t st) {
    if (B > 0) advance_streams_kernel<<<(B + 127) / 128, 128, 0, st>>>(slot_of_b, B, T, ring_pos, valid_len);
}

}  // namespace nsb
