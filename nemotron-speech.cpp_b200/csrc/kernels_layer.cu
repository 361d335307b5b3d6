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
#include <cstdlib>

#include "kernels.cuh"

namespace nsb {

NSB_DEFINE_TRACE_BINDER(trace_bind_layer)

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

// Mean and (biased) variance of the 1024 values a block of 256 threads holds 4 apiece, in ONE block-wide reduction: Chan's
// pairwise merge of (mean, M2) for equal group sizes -- as accurate as the two-pass form (no E[x^2] - mean^2 cancellation) at
// half the barriers. Used by the 16-bit engine modes; strict fp32 keeps the reference's two-pass order.
__device__ __forceinline__ void block_mean_var_256(float v0, float v1, float v2, float v3, float* red /*[16]*/, float& mean, float& var) {
    float m = ((v0 + v1) + (v2 + v3)) * 0.25f;
    float M2 = (v0 - m) * (v0 - m) + (v1 - m) * (v1 - m) + (v2 - m) * (v2 - m) + (v3 - m) * (v3 - m);
    float half_n = 2.0f;                                                          // n / 2 of each of the two groups being merged
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float mo = __shfl_xor_sync(0xffffffffu, m, o), M2o = __shfl_xor_sync(0xffffffffu, M2, o);
        const float d = mo - m;
        M2 = (M2 + M2o) + d * d * half_n; m = m + 0.5f * d; half_n *= 2.0f;
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();                                                              // protect red[] reuse
    if (l == 0) { red[2 * w] = m; red[2 * w + 1] = M2; }
    __syncthreads();
    float mm[8], MM[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { mm[i] = red[2 * i]; MM[i] = red[2 * i + 1]; }
#pragma unroll
    for (int n = 8; n > 1; n >>= 1) {                                             // 8 warps of 128 values -> 4 x 256 -> 2 x 512 -> 1024
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float d = mm[2 * i + 1] - mm[2 * i];
            MM[i] = (MM[2 * i] + MM[2 * i + 1]) + d * d * half_n; mm[i] = mm[2 * i] + 0.5f * d;
        }
        half_n *= 2.0f;
    }
    mean = mm[0]; var = MM[0] * (1.0f / D_MODEL);
}

// LayerNorm over rows of 1024, eps 1e-5, two-pass (mean, then variance of centred values) like ggml_norm.
// One CTA (256 threads x 4 channels) per row.
__device__ __forceinline__ float4 load_x_reduced(float* x, size_t o, const PartialSum& ps, int rows) {
    float4 v = *(const float4*)(x + o);
    if (ps.n > 0) {                                                              // fold the split-K partials of the previous GEMM
        float4 t[8];                                                             // all loads in flight at once; summed in slice order
#pragma unroll
        for (int s = 0; s < 8; ++s) if (s < ps.n) t[s] = __ldcg((const float4*)(ps.part + (size_t)s * rows * D_MODEL + o));
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int s = 0; s < 8; ++s) if (s < ps.n) { acc.x += t[s].x; acc.y += t[s].y; acc.z += t[s].z; acc.w += t[s].w; }
        v.x += ps.alpha * acc.x; v.y += ps.alpha * acc.y; v.z += ps.alpha * acc.z; v.w += ps.alpha * acc.w;
        *(float4*)(x + o) = v;
    }
    return v;
}

__global__ void __launch_bounds__(256) layernorm_kernel(float* x, const float* __restrict__ g,
                                                        const float* __restrict__ b, void* y, int out_type, const PartialSum ps) {
    NSB_KERNEL_BEGIN(TR_LN)
    __shared__ float red[16];
    const int row = blockIdx.x, c = threadIdx.x * 4;
    const float4 gg = *(const float4*)(g + c), bb = *(const float4*)(b + c);    // static: before the dependency wait
    NSB_KERNEL_WAIT()
    const float4 v = load_x_reduced(x, (size_t)row * D_MODEL + c, ps, gridDim.x);
    float mean, var;
    if (out_type == OUT_F32) {
        mean = block_sum_256(v.x + v.y + v.z + v.w, red) * (1.0f / D_MODEL);
        const float e0 = v.x - mean, e1 = v.y - mean, e2 = v.z - mean, e3 = v.w - mean;
        var = block_sum_256(e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3, red) * (1.0f / D_MODEL);
    } else block_mean_var_256(v.x, v.y, v.z, v.w, red, mean, var);
    const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
    const float rs = 1.0f / sqrtf(var + 1e-5f);
    const size_t o = (size_t)row * D_MODEL + c;
    const float o0 = d0 * rs * gg.x + bb.x, o1 = d1 * rs * gg.y + bb.y, o2 = d2 * rs * gg.z + bb.z, o3 = d3 * rs * gg.w + bb.w;
    if (out_type == OUT_F32) *(float4*)((float*)y + o) = make_float4(o0, o1, o2, o3);
    else if (out_type == OUT_F16) { __half2* p = (__half2*)((__half*)y + o); p[0] = __floats2half2_rn(o0, o1); p[1] = __floats2half2_rn(o2, o3); }
    else { __nv_bfloat162* p = (__nv_bfloat162*)((__nv_bfloat16*)y + o); p[0] = __floats2bfloat162_rn(o0, o1); p[1] = __floats2bfloat162_rn(o2, o3); }
    NSB_KERNEL_EPILOGUE();
}

// ------------------------------------------------------------------------------------------
// Large batches (>= 1024 rows, 16-bit modes; measured neutral at 896 rows, +2 % on the step at 1792): one WARP per row instead of one CTA per row. A lane owns 8 float4 of the row (all eight
// loads -- and those of each split-K plane -- in flight at once), statistics by the same pairwise (mean, M2) merge through five
// shuffles, no shared memory, no block barrier; 8 rows per CTA. At 1792 rows the one-CTA-per-row kernels spend most of their time in
// two block-wide barriers per row with 1792 CTAs in flight; this form is bound by its L2 traffic. g2 != nullptr: the fused pair
// norm_out -> norm_feed_forward1 of the next layer (layernorm2_kernel); y == nullptr with g2 == nullptr: norm_out of the last layer.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_mean_var_1024(const float4 (&v)[8], float& mean, float& var) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    float m = s * (1.0f / 32.0f), M2 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { const float a = v[j].x - m, b = v[j].y - m, c = v[j].z - m, d = v[j].w - m; M2 += (a * a + b * b) + (c * c + d * d); }
    float half_n = 16.0f;                                                         // n / 2 of each of the two groups being merged
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float mo = __shfl_xor_sync(0xffffffffu, m, o), M2o = __shfl_xor_sync(0xffffffffu, M2, o);
        const float d = mo - m;
        M2 = (M2 + M2o) + d * d * half_n; m = m + 0.5f * d; half_n *= 2.0f;
    }
    mean = m; var = M2 * (1.0f / D_MODEL);
}
__device__ __forceinline__ void store4_out(void* y, size_t o, float4 r, int out_type) {
    if (out_type == OUT_F32) *(float4*)((float*)y + o) = r;
    else if (out_type == OUT_F16) { __half2* p = (__half2*)((__half*)y + o); p[0] = __floats2half2_rn(r.x, r.y); p[1] = __floats2half2_rn(r.z, r.w); }
    else { __nv_bfloat162* p = (__nv_bfloat162*)((__nv_bfloat16*)y + o); p[0] = __floats2bfloat162_rn(r.x, r.y); p[1] = __floats2bfloat162_rn(r.z, r.w); }
}

__global__ void __launch_bounds__(256) layernorm_rows_kernel(float* x, int rows, const float* __restrict__ g1, const float* __restrict__ b1,
                                                             const float* __restrict__ g2, const float* __restrict__ b2, void* y, int out_type,
                                                             const PartialSum ps) {
    NSB_KERNEL_BEGIN(g2 ? TR_LN2 : TR_LN)
    const int lane = threadIdx.x & 31, row = blockIdx.x * 8 + (threadIdx.x >> 5);
    NSB_KERNEL_WAIT()
    if (row >= rows) return;                                                      // warp-uniform
    float* xr = x + (size_t)row * D_MODEL;
    float4 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = *(const float4*)(xr + 4 * (32 * j + lane));
    if (ps.n > 0) {                                                               // fold the split-K partials of the previous GEMM: x += alpha * (p0 + p1 + ...), slice order
        float4 acc[8];
        for (int sl = 0; sl < ps.n; ++sl) {
            const float* pr = ps.part + ((size_t)sl * rows + row) * D_MODEL;
            float4 t[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) t[j] = __ldcg((const float4*)(pr + 4 * (32 * j + lane)));
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (sl == 0) { acc[j] = make_float4(0.f + t[j].x, 0.f + t[j].y, 0.f + t[j].z, 0.f + t[j].w); }
                else { acc[j].x += t[j].x; acc[j].y += t[j].y; acc[j].z += t[j].z; acc[j].w += t[j].w; }
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) { v[j].x += ps.alpha * acc[j].x; v[j].y += ps.alpha * acc[j].y; v[j].z += ps.alpha * acc[j].z; v[j].w += ps.alpha * acc[j].w; }
        if (!g2 && y) {                                                           // plain LayerNorm: x keeps the reduced residual stream
#pragma unroll
            for (int j = 0; j < 8; ++j) *(float4*)(xr + 4 * (32 * j + lane)) = v[j];
        }
    }
    float mean, var;
    warp_mean_var_1024(v, mean, var);
    float rs = 1.0f / sqrtf(var + 1e-5f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 gg = *(const float4*)(g1 + 4 * (32 * j + lane)), bb = *(const float4*)(b1 + 4 * (32 * j + lane));
        v[j] = make_float4((v[j].x - mean) * rs * gg.x + bb.x, (v[j].y - mean) * rs * gg.y + bb.y, (v[j].z - mean) * rs * gg.z + bb.z, (v[j].w - mean) * rs * gg.w + bb.w);
    }
    if (!g2 && y) {                                                               // LN(x) -> y (operand of the next GEMM)
#pragma unroll
        for (int j = 0; j < 8; ++j) store4_out(y, (size_t)row * D_MODEL + 4 * (32 * j + lane), v[j], out_type);
        NSB_KERNEL_EPILOGUE();
        return;
    }
    // norm_out: x <- LN1(x) (f32, in place) ...
#pragma unroll
    for (int j = 0; j < 8; ++j) *(float4*)(xr + 4 * (32 * j + lane)) = v[j];
    if (g2) {                                                                     // ... and y <- LN2(x) for the next layer's first GEMM
        warp_mean_var_1024(v, mean, var);
        rs = 1.0f / sqrtf(var + 1e-5f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 gg = *(const float4*)(g2 + 4 * (32 * j + lane)), bb = *(const float4*)(b2 + 4 * (32 * j + lane));
            const float4 r = make_float4((v[j].x - mean) * rs * gg.x + bb.x, (v[j].y - mean) * rs * gg.y + bb.y, (v[j].z - mean) * rs * gg.z + bb.z, (v[j].w - mean) * rs * gg.w + bb.w);
            store4_out(y, (size_t)row * D_MODEL + 4 * (32 * j + lane), r, out_type);
        }
    }
    NSB_KERNEL_EPILOGUE();
}
static bool ln_rows_enabled(int rows, int out_type) {
    static const int min_rows = [] { const char* e = getenv("NSB_LN_ROWS_MIN"); return e ? atoi(e) : 1024; }();
    return out_type != OUT_F32 && rows >= min_rows;                               // strict fp32 keeps the two-pass block kernels
}
void launch_layernorm(float* x, int rows, const float* g, const float* b, void* y, int out_type, const PartialSum& ps, cudaStream_t st) {
    if (rows > 0 && ln_rows_enabled(rows, out_type)) {
        launch_k(layernorm_rows_kernel, dim3((rows + 7) / 8), dim3(256), 0, st, x, rows, g, b, (const float*)nullptr, (const float*)nullptr, y, out_type, ps);
        return;
    }
    if (rows > 0) launch_k(layernorm_kernel, dim3(rows), dim3(256), 0, st, x, g, b, y, out_type, ps);
}

// norm_out of layer l fused with norm_feed_forward1 of layer l+1: x <- LN1(x) (f32, in place); y2 <- LN2(x)
__global__ void __launch_bounds__(256) layernorm2_kernel(float* x, const float* __restrict__ g1, const float* __restrict__ b1,
                                                         const float* __restrict__ g2, const float* __restrict__ b2,
                                                         void* y2, int out_type, const PartialSum ps) {
    NSB_KERNEL_BEGIN(TR_LN2)
    __shared__ float red[16];
    const int row = blockIdx.x, c = threadIdx.x * 4;
    const size_t o = (size_t)row * D_MODEL + c;
    const bool one_pass = out_type != OUT_F32;
    float4 gg = *(const float4*)(g1 + c), bb = *(const float4*)(b1 + c);       // static: before the dependency wait
    float4 gg2 = make_float4(0.f, 0.f, 0.f, 0.f), bb2 = gg2;
    if (y2 != nullptr) { gg2 = *(const float4*)(g2 + c); bb2 = *(const float4*)(b2 + c); }
    NSB_KERNEL_WAIT()
    float4 v = load_x_reduced(x, o, ps, gridDim.x);
    float mean, var, d0, d1, d2, d3;
    auto stats = [&]() {
        if (one_pass) block_mean_var_256(v.x, v.y, v.z, v.w, red, mean, var);
        else {
            mean = block_sum_256(v.x + v.y + v.z + v.w, red) * (1.0f / D_MODEL);
            const float e0 = v.x - mean, e1 = v.y - mean, e2 = v.z - mean, e3 = v.w - mean;
            var = block_sum_256(e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3, red) * (1.0f / D_MODEL);
        }
        d0 = v.x - mean; d1 = v.y - mean; d2 = v.z - mean; d3 = v.w - mean;
    };
    stats();
    float rs = 1.0f / sqrtf(var + 1e-5f);
    v = make_float4(d0 * rs * gg.x + bb.x, d1 * rs * gg.y + bb.y, d2 * rs * gg.z + bb.z, d3 * rs * gg.w + bb.w);
    *(float4*)(x + o) = v;
    if (y2 == nullptr) { NSB_KERNEL_EPILOGUE(); return; }                        // last layer: no following norm (uniform branch)
    stats();
    rs = 1.0f / sqrtf(var + 1e-5f);
    gg = gg2; bb = bb2;
    const float o0 = d0 * rs * gg.x + bb.x, o1 = d1 * rs * gg.y + bb.y, o2 = d2 * rs * gg.z + bb.z, o3 = d3 * rs * gg.w + bb.w;
    if (out_type == OUT_F32) *(float4*)((float*)y2 + o) = make_float4(o0, o1, o2, o3);
    else if (out_type == OUT_F16) { __half2* p = (__half2*)((__half*)y2 + o); p[0] = __floats2half2_rn(o0, o1); p[1] = __floats2half2_rn(o2, o3); }
    else { __nv_bfloat162* p = (__nv_bfloat162*)((__nv_bfloat16*)y2 + o); p[0] = __floats2bfloat162_rn(o0, o1); p[1] = __floats2bfloat162_rn(o2, o3); }
    NSB_KERNEL_EPILOGUE();
}
void launch_layernorm2(float* x, int rows, const float* g1, const float* b1, const float* g2, const float* b2, void* y2, int out_type,
                       const PartialSum& ps, cudaStream_t st) {
    if (rows > 0 && ln_rows_enabled(rows, out_type)) {
        launch_k(layernorm_rows_kernel, dim3((rows + 7) / 8), dim3(256), 0, st, x, rows, g1, b1, g2, b2, y2, out_type, ps);
        return;
    }
    if (rows > 0) launch_k(layernorm2_kernel, dim3(rows), dim3(256), 0, st, x, g1, b1, g2, b2, y2, out_type, ps);
}

// ------------------------------------------------------------------------------------------
// Cached relative-position attention. One CTA per (head, batch row); 256 threads = 8 warps.
//   keys j = 0..K-1 (K = L+T): j < L are ring rows (oldest first), j >= L are this chunk's new rows
//   score[i][j] = ((q_i + u) . k_j + (q_i + v) . P[L + i - j]) / sqrt(128),  j >= L - valid_len
//   ctx[i] = softmax_j(score[i]) . v_j
// One memory round trip per CTA: the valid ring rows of K and V for this head (the only HBM traffic, each byte read
// exactly once) are fetched with a single burst of 16-byte cp.async into shared memory; while they are in flight the
// warps compute the positional term (q+v).P[r] straight from the L2-resident projected table (one row per warp step,
// all rows of a warp requested before the first is used) and append this chunk's K/V rows to the ring (positions
// (w + i) mod (L+T), replacing the reference's concat + roll, nemo-stream.cpp:465-484). Lanes hold 4 of the 128 head dims
// of (q+u), (q+v) for up to TQ queries in registers. 16-bit K/V keep the CTA at ~44 KB of shared memory, so all
// 8 x 64 CTAs of a 64-stream step are resident at once (4 per SM).
// ------------------------------------------------------------------------------------------
template <int KV> struct KvT { using type = float; };
template <> struct KvT<1> { using type = __half; };
template <> struct KvT<2> { using type = __nv_bfloat16; };

constexpr int ATT_MAX_T = 65, ATT_MAX_K = ATT_L + ATT_MAX_T + 1, ATT_MAX_REL = ATT_L + 2 * ATT_MAX_T;
constexpr int ATT_PROWS = 5;                                     // positional rows per warp held in flight

__device__ __forceinline__ void load4(const float* p, float (&f)[4]) {
    const float4 v = *reinterpret_cast<const float4*>(p); f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
}
__device__ __forceinline__ void load4(const __half* p, float (&f)[4]) {
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
}
__device__ __forceinline__ void load4(const __nv_bfloat16* p, float (&f)[4]) {
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.x)), b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.y));
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
}
__device__ __forceinline__ void load2(const float* p, float (&f)[2]) { const float2 v = *reinterpret_cast<const float2*>(p); f[0] = v.x; f[1] = v.y; }
__device__ __forceinline__ void load2(const __half* p, float (&f)[2]) { const float2 v = __half22float2(*reinterpret_cast<const __half2*>(p)); f[0] = v.x; f[1] = v.y; }
__device__ __forceinline__ void load2(const __nv_bfloat16* p, float (&f)[2]) { const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p)); f[0] = v.x; f[1] = v.y; }
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}

// f32 vector load that folds up to MAXP split-K partial planes (fixed order z = 0, 1, ...: deterministic), all loads issued first
template <int MAXP>
__device__ __forceinline__ float4 ld4_planes(const float* p, int planes, long long stride) {
    float4 v = *reinterpret_cast<const float4*>(p);
    if (planes > 1) {
        float4 t[MAXP - 1];
#pragma unroll
        for (int z = 1; z < MAXP; ++z) if (z < planes) t[z - 1] = *reinterpret_cast<const float4*>(p + (size_t)z * stride);
#pragma unroll
        for (int z = 1; z < MAXP; ++z) if (z < planes) { v.x += t[z - 1].x; v.y += t[z - 1].y; v.z += t[z - 1].z; v.w += t[z - 1].w; }
    }
    return v;
}
constexpr int QKV_MAX_PLANES = 2, PW1_MAX_PLANES = 4;
__device__ __forceinline__ void load4_planes(const float* p, int planes, long long stride, float (&f)[4]) {
    const float4 v = ld4_planes<QKV_MAX_PLANES>(p, planes, stride); f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
}

// Sum NV per-lane partial values over the warp by recursive halving: at every exchange the lanes split the values between
// them, so NV sums cost (NV - 1) + (5 - log2 NV) shuffles instead of 5 NV. On return lane L holds the complete sum of value
// `idx` (its top log2 NV lane bits, MSB first); lanes that differ only in the lower bits hold copies -- `owner` marks one.
template <int NV>
__device__ __forceinline__ float warp_sum_halving(float (&s)[NV], int lane, int& idx, bool& owner) {
    static_assert(NV == 1 || NV == 2 || NV == 4 || NV == 8 || NV == 16, "power of two");
    int n = NV, bit = 16;
    idx = 0;
#pragma unroll
    for (int step = 0; step < 5; ++step) {
        if (n > 1) {
            const bool hi = (lane & bit) != 0;
            n >>= 1;
#pragma unroll
            for (int k = 0; k < NV / 2; ++k) {
                if (k < n) {
                    const float send = hi ? s[k] : s[k + n], keep = hi ? s[k + n] : s[k];
                    s[k] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
                }
            }
            idx = idx * 2 + (hi ? 1 : 0);
        } else {
            s[0] += __shfl_xor_sync(0xffffffffu, s[0], bit);
        }
        bit >>= 1;
    }
    owner = (lane & (32 / NV - 1)) == 0;
    return s[0];
}

template <int TQ> struct AttnSmemF {                             // fp32 part of the shared memory (K / V tiles follow)
    float ac[TQ][ATT_MAX_K];                                     // (q+u).k, then probabilities
    float bd[TQ][ATT_MAX_REL];                                   // (q+v).P[r]
    float red[4][TQ][D_HEAD];
};

template <int KV, int TQ>
__global__ void __launch_bounds__(256, TQ <= 2 ? 4 : 2) attention_kernel(const AttnArgs a) {
    using E = typename KvT<KV>::type;
    constexpr int EPV = 16 / sizeof(E), VPR = D_HEAD / EPV;                     // elements per 16-byte vector, vectors per row
    extern __shared__ __align__(16) uint8_t att_smem[];
    AttnSmemF<TQ>& sf = *reinterpret_cast<AttnSmemF<TQ>*>(att_smem);
    const int T = a.T, K = ATT_L + T, Cap = K;
    E* Ks = reinterpret_cast<E*>(att_smem + sizeof(AttnSmemF<TQ>));             // [K][128]
    E* Vs = Ks + (size_t)K * D_HEAD;                                            // [K][128]
    const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // Before the dependency wait: everything that does not depend on this step's QKV projection -- the cached K / V rows
    // (written by this layer's attention kernel of EARLIER steps only), ring bookkeeping, the positional biases.
    NSB_KERNEL_BEGIN(TR_ATTN)
    const int slot = a.slot_of_b[b], w = a.ring_pos[slot], valid = a.valid_len[slot];
    const int first = ATT_L - valid;                                           // keys j < first are not yet valid (:982-992)
    const float* qkv = a.qkv + (size_t)b * T * 3 * D_MODEL + h * D_HEAD;
    E* kring = reinterpret_cast<E*>(a.k_ring) + (size_t)slot * a.slot_stride + h * D_HEAD;
    E* vring = reinterpret_cast<E*>(a.v_ring) + (size_t)slot * a.slot_stride + h * D_HEAD;
    const E* P = reinterpret_cast<const E*>(a.pos_proj) + h * D_HEAD;           // L2-resident (146 KB per layer in 16-bit modes)
    auto ring_row = [&](int j) { return (size_t)((w + Cap - ATT_L + j) % Cap) * D_MODEL; };

    // ---- one burst: valid cached K / V rows -> shared memory ----
    for (int e = tid; e < (ATT_L - first) * VPR; e += 256) {
        const int j = first + e / VPR, c = (e % VPR) * EPV;
        const size_t g = ring_row(j) + c;
        cp_async16(Ks + (size_t)j * D_HEAD + c, kring + g);
        cp_async16(Vs + (size_t)j * D_HEAD + c, vring + g);
    }
    const float scale = 0.08838834764831845f;                                   // 1/sqrt(128) :517
    const int c4 = lane * 4;
    float bu[4], bv[4];
    load4(a.bias_u + h * D_HEAD + c4, bu); load4(a.bias_v + h * D_HEAD + c4, bv);
    NSB_KERNEL_WAIT()
    // ---- this chunk's K / V rows (rounded to the ring dtype): shared memory AND ring append ----
    for (int e = tid; e < T * (D_HEAD / 4); e += 256) {
        const int i = e / (D_HEAD / 4), c = (e % (D_HEAD / 4)) * 4;
        const float4 kn = ld4_planes<QKV_MAX_PLANES>(qkv + (size_t)i * 3 * D_MODEL + D_MODEL + c, a.planes, a.plane_stride);
        const float4 vn = ld4_planes<QKV_MAX_PLANES>(qkv + (size_t)i * 3 * D_MODEL + 2 * D_MODEL + c, a.planes, a.plane_stride);
        const E k4[4] = {from_f32<E>(kn.x), from_f32<E>(kn.y), from_f32<E>(kn.z), from_f32<E>(kn.w)};
        const E v4[4] = {from_f32<E>(vn.x), from_f32<E>(vn.y), from_f32<E>(vn.z), from_f32<E>(vn.w)};
        E* kd = kring + ring_row(ATT_L + i) + c; E* vd = vring + ring_row(ATT_L + i) + c;
        E* ks = Ks + (size_t)(ATT_L + i) * D_HEAD + c; E* vs = Vs + (size_t)(ATT_L + i) * D_HEAD + c;
#pragma unroll
        for (int u = 0; u < 4; ++u) { kd[u] = k4[u]; vd[u] = v4[u]; ks[u] = k4[u]; vs[u] = v4[u]; }
    }

    for (int q0 = 0; q0 < T; q0 += TQ) {
        const int nq = min(TQ, T - q0);
        float qu[TQ][4], qv[TQ][4];
#pragma unroll
        for (int i = 0; i < TQ; ++i) {
            float q[4] = {0.f, 0.f, 0.f, 0.f};
            if (i < nq) load4_planes(qkv + (size_t)(q0 + i) * 3 * D_MODEL + c4, a.planes, a.plane_stride, q);
#pragma unroll
            for (int u = 0; u < 4; ++u) { qu[i][u] = q[u] + bu[u]; qv[i][u] = q[u] + bv[u]; }   // :503-507
        }
        // ---- BD_raw[i][r] = (q_i + v) . P[r], r = rel + (T-1), from L2 (ATT_PROWS rows per warp in flight) ----
        const int r_end = ATT_L + T - 1 + (q0 + nq - 1) - first + 1;            // largest row index used + 1
        for (int rb = 0; rb < r_end; rb += 8 * ATT_PROWS) {
            float pf[ATT_PROWS][4];
#pragma unroll
            for (int r = 0; r < ATT_PROWS; ++r) { const int rr = rb + warp + 8 * r; if (rr < r_end) load4(P + (size_t)rr * D_MODEL + c4, pf[r]); }
#pragma unroll
            for (int r = 0; r < ATT_PROWS; ++r) {
                const int rr = rb + warp + 8 * r;
                if (rr < r_end) {
                    if constexpr (KV == 0) {                                      // strict fp32 mode keeps the original summation order
#pragma unroll
                        for (int i = 0; i < TQ; ++i) {
                            float s = qv[i][0] * pf[r][0];
                            s = fmaf(qv[i][1], pf[r][1], s); s = fmaf(qv[i][2], pf[r][2], s); s = fmaf(qv[i][3], pf[r][3], s);
                            s = warp_sum(s);
                            if (lane == 0) sf.bd[i][rr] = s;
                        }
                    } else {                                                      // TQ sums per row by recursive halving (the kernel was shuffle-issue bound)
                        float sv[TQ];
#pragma unroll
                        for (int i = 0; i < TQ; ++i) {
                            float s = qv[i][0] * pf[r][0];
                            s = fmaf(qv[i][1], pf[r][1], s); s = fmaf(qv[i][2], pf[r][2], s); sv[i] = fmaf(qv[i][3], pf[r][3], s);
                        }
                        int idx; bool owner;
                        const float v = warp_sum_halving<TQ>(sv, lane, idx, owner);
                        if (owner) sf.bd[idx][rr] = v;
                    }
                }
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                                                          // K / V tiles, new rows and BD visible to all warps
        if (tr_slot >= 0 && q0 == 0) trace_mark(tr_slot, 3);                      // trace: BD + K/V landed
        // ---- AC[i][j] = (q_i + u) . k_j : one key row per warp step out of shared memory ----
        for (int j = first + warp; j < K; j += 8) {
            float kf[4];
            load4(Ks + (size_t)j * D_HEAD + c4, kf);
            if constexpr (KV == 0) {
#pragma unroll
                for (int i = 0; i < TQ; ++i) {
                    float s = qu[i][0] * kf[0];
                    s = fmaf(qu[i][1], kf[1], s); s = fmaf(qu[i][2], kf[2], s); s = fmaf(qu[i][3], kf[3], s);
                    s = warp_sum(s);
                    if (lane == 0) sf.ac[i][j] = s;
                }
            } else {
                float sv[TQ];
#pragma unroll
                for (int i = 0; i < TQ; ++i) {
                    float s = qu[i][0] * kf[0];
                    s = fmaf(qu[i][1], kf[1], s); s = fmaf(qu[i][2], kf[2], s); sv[i] = fmaf(qu[i][3], kf[3], s);
                }
                int idx; bool owner;
                const float v = warp_sum_halving<TQ>(sv, lane, idx, owner);
                if (owner) sf.ac[idx][j] = v;
            }
        }
        __syncthreads();
        // ---- softmax over valid keys; rel-shift = index arithmetic: BD[i][j] = BD_raw[i][L + i - j + T-1] (:391-433) ----
        for (int i = warp; i < nq; i += 8) {
            const int qi = q0 + i;
            float mx = -INFINITY;
            for (int j = first + lane; j < K; j += 32) {
                const float s = (sf.ac[i][j] + sf.bd[i][ATT_L + qi - j + T - 1]) * scale;
                sf.ac[i][j] = s; mx = fmaxf(mx, s);
            }
            mx = warp_max(mx);
            float sum = 0.f;
            for (int j = first + lane; j < K; j += 32) { const float e = expf(sf.ac[i][j] - mx); sf.ac[i][j] = e; sum += e; }
            sum = warp_sum(sum);
            const float inv = 1.0f / sum;
            for (int j = first + lane; j < K; j += 32) sf.ac[i][j] *= inv;
        }
        __syncthreads();
        if (tr_slot >= 0 && q0 == 0) trace_mark(tr_slot, 4);                      // trace: AC + softmax done
        // ---- ctx[i][d] = sum_j p[i][j] v_j[d]; thread = (key group kg, 2 dims) ----
        {
            const int d2 = (tid & 63) * 2, kg = tid >> 6;
            float acc[TQ][2];
#pragma unroll
            for (int i = 0; i < TQ; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; }
#pragma unroll 4
            for (int j = first + kg; j < K; j += 4) {
                float vf[2];
                load2(Vs + (size_t)j * D_HEAD + d2, vf);
#pragma unroll
                for (int i = 0; i < TQ; ++i) {
                    const float p = sf.ac[i][j];
                    acc[i][0] = fmaf(p, vf[0], acc[i][0]); acc[i][1] = fmaf(p, vf[1], acc[i][1]);
                }
            }
#pragma unroll
            for (int i = 0; i < TQ; ++i) { sf.red[kg][i][d2] = acc[i][0]; sf.red[kg][i][d2 + 1] = acc[i][1]; }
        }
        __syncthreads();
        for (int e = tid; e < nq * D_HEAD; e += 256) {
            const int i = e / D_HEAD, d = e % D_HEAD;
            const float v = (sf.red[0][i][d] + sf.red[1][i][d]) + (sf.red[2][i][d] + sf.red[3][i][d]);
            store_out(a.ctx, ((size_t)b * T + q0 + i) * D_MODEL + h * D_HEAD + d, v, a.out_type);
        }
        __syncthreads();                                                          // ac / bd / red are reused by the next query tile
    }
    NSB_KERNEL_EPILOGUE();
}

// ------------------------------------------------------------------------------------------
// Low-latency variant for T <= 2 with a 16-bit ring (the 80 / 160 ms modes): one CTA = one head x TWO streams, 512 threads
// (each half of the CTA is the kernel above for one stream, synchronised by its own named barrier). What the pairing buys:
// the positional rows P[r] of the head (73 x 128, the same for every stream) are staged ONCE per CTA in shared memory, by
// cp.async issued BEFORE the dependency wait together with the K / V ring rows -- so after the QKV GEMM completes nothing
// but q and this chunk's k / v rows (L2-hot, one round trip) stands between the wait and the arithmetic. In the one-stream
// kernel the (q+v).P products wait for two dependent L2 round trips per warp AFTER the wait (7.4 of its 11 us, in-graph trace).
// 104 KB of shared memory per CTA -> 2 CTAs per SM -> all 8 x B/2 CTAs of a 64-stream step in one wave.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void half_barrier(int half) { asm volatile("bar.sync %0, 256;" ::"r"(1 + half) : "memory"); }

template <int KV, int TQ>
__global__ void __launch_bounds__(512, 2) attention_pair_kernel(const AttnArgs a) {
    using E = typename KvT<KV>::type;
    constexpr int EPV = 16 / sizeof(E), VPR = D_HEAD / EPV;
    extern __shared__ __align__(16) uint8_t att_smem[];
    const int T = a.T, K = ATT_L + T, Cap = K, n_rel = ATT_L + 2 * T - 1;
    const int tid = threadIdx.x, half = tid >> 8, t = tid & 255, warp = t >> 5, lane = t & 31;
    const int h = blockIdx.x, b = blockIdx.y * 2 + half;
    const bool active = b < a.B;
    E* Ps = reinterpret_cast<E*>(att_smem);                                     // [n_rel][128], shared by both halves
    uint8_t* hb = att_smem + (size_t)n_rel * D_HEAD * sizeof(E) + (size_t)half * (sizeof(AttnSmemF<TQ>) + (size_t)2 * K * D_HEAD * sizeof(E));
    AttnSmemF<TQ>& sf = *reinterpret_cast<AttnSmemF<TQ>*>(hb);
    E* Ks = reinterpret_cast<E*>(hb + sizeof(AttnSmemF<TQ>));                   // [K][128]
    E* Vs = Ks + (size_t)K * D_HEAD;                                            // [K][128]
    NSB_KERNEL_BEGIN(TR_ATTN)
    // ---- before the dependency wait: positional rows, cached K / V rows, ring bookkeeping, biases ----
    const E* P = reinterpret_cast<const E*>(a.pos_proj) + h * D_HEAD;
    for (int e = tid; e < n_rel * VPR; e += 512) {
        const int r = e / VPR, c = (e % VPR) * EPV;
        cp_async16(Ps + (size_t)r * D_HEAD + c, P + (size_t)r * D_MODEL + c);
    }
    int slot = 0, w = 0, first = ATT_L;
    if (active) { slot = a.slot_of_b[b]; w = a.ring_pos[slot]; first = ATT_L - a.valid_len[slot]; }   // keys j < first are not yet valid (:982-992)
    const float* qkv = a.qkv + (size_t)b * T * 3 * D_MODEL + h * D_HEAD;
    E* kring = reinterpret_cast<E*>(a.k_ring) + (size_t)slot * a.slot_stride + h * D_HEAD;
    E* vring = reinterpret_cast<E*>(a.v_ring) + (size_t)slot * a.slot_stride + h * D_HEAD;
    auto ring_row = [&](int j) { return (size_t)((w + Cap - ATT_L + j) % Cap) * D_MODEL; };
    for (int e = t; e < (ATT_L - first) * VPR; e += 256) {
        const int j = first + e / VPR, c = (e % VPR) * EPV;
        const size_t g = ring_row(j) + c;
        cp_async16(Ks + (size_t)j * D_HEAD + c, kring + g);
        cp_async16(Vs + (size_t)j * D_HEAD + c, vring + g);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    const float scale = 0.08838834764831845f;                                   // 1/sqrt(128) :517
    const int c4 = lane * 4;
    float bu[4], bv[4];
    load4(a.bias_u + h * D_HEAD + c4, bu); load4(a.bias_v + h * D_HEAD + c4, bv);
    NSB_KERNEL_WAIT()
    float qu[TQ][4], qv[TQ][4];
    if (active) {
        // q of every query, and this chunk's K / V rows (rounded to the ring dtype): shared memory AND ring append
#pragma unroll
        for (int i = 0; i < TQ; ++i) {
            float q[4] = {0.f, 0.f, 0.f, 0.f};
            if (i < T) load4_planes(qkv + (size_t)i * 3 * D_MODEL + c4, a.planes, a.plane_stride, q);
#pragma unroll
            for (int u = 0; u < 4; ++u) { qu[i][u] = q[u] + bu[u]; qv[i][u] = q[u] + bv[u]; }   // :503-507
        }
        for (int e = t; e < T * (D_HEAD / 4); e += 256) {
            const int i = e / (D_HEAD / 4), c = (e % (D_HEAD / 4)) * 4;
            const float4 kn = ld4_planes<QKV_MAX_PLANES>(qkv + (size_t)i * 3 * D_MODEL + D_MODEL + c, a.planes, a.plane_stride);
            const float4 vn = ld4_planes<QKV_MAX_PLANES>(qkv + (size_t)i * 3 * D_MODEL + 2 * D_MODEL + c, a.planes, a.plane_stride);
            const E k4[4] = {from_f32<E>(kn.x), from_f32<E>(kn.y), from_f32<E>(kn.z), from_f32<E>(kn.w)};
            const E v4[4] = {from_f32<E>(vn.x), from_f32<E>(vn.y), from_f32<E>(vn.z), from_f32<E>(vn.w)};
            E* kd = kring + ring_row(ATT_L + i) + c; E* vd = vring + ring_row(ATT_L + i) + c;
            E* ks = Ks + (size_t)(ATT_L + i) * D_HEAD + c; E* vs = Vs + (size_t)(ATT_L + i) * D_HEAD + c;
#pragma unroll
            for (int u = 0; u < 4; ++u) { kd[u] = k4[u]; vd[u] = v4[u]; ks[u] = k4[u]; vs[u] = v4[u]; }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                                              // the only CTA-wide barrier: P / K / V tiles and the new rows are in place
    if (tr_slot >= 0) trace_mark(tr_slot, 3);
    if (!active) return;                                                          // odd batch: the second half of the last CTA has no stream
    // ---- BD_raw[i][r] = (q_i + v) . P[r] and AC[i][j] = (q_i + u) . k_j: one row per warp step, out of shared memory ----
    // Four rows per warp step. The 4 x TQ per-lane partial dot products are summed over the warp by recursive halving (the
    // lanes split the values between them at every exchange): 9 shuffles for 8 sums (6 for 4) instead of 5 per sum. With 32
    // warps per SM doing nothing but these reductions the kernel was bound by shuffle issue (one warp shuffle per clock per SM).
    const int r_end = ATT_L + 2 * T - 1 - first;                                  // largest positional row used + 1
    auto dots4 = [&](const E* rows, int r0, int r_lim, const float (&qq)[TQ][4], float* out, int out_stride) {
        constexpr int NV = 4 * TQ;                                                // value index v = u * TQ + i (row r0 + 8 u, query i)
        float s[NV];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int rr = r0 + 8 * u;
            float f[4] = {0.f, 0.f, 0.f, 0.f};
            if (rr < r_lim) load4(rows + (size_t)rr * D_HEAD + c4, f);
#pragma unroll
            for (int i = 0; i < TQ; ++i) {
                float v = qq[i][0] * f[0];
                v = fmaf(qq[i][1], f[1], v); v = fmaf(qq[i][2], f[2], v); v = fmaf(qq[i][3], f[3], v);
                s[u * TQ + i] = v;
            }
        }
        // halving steps: after the exchange over lane bit `bit`, a lane keeps the upper half of the values if that bit is set
        int idx = 0;                                                              // index of the value this lane ends up owning
        int n = NV, bit = 16;
#pragma unroll
        for (int step = 0; step < (TQ == 2 ? 3 : 2); ++step) {
            const bool hi = (lane & bit) != 0;
            n >>= 1;
#pragma unroll
            for (int k = 0; k < NV / 2; ++k) {
                if (k < n) {
                    const float send = hi ? s[k] : s[k + n], keep = hi ? s[k + n] : s[k];
                    s[k] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
                }
            }
            idx = idx * 2 + (hi ? 1 : 0);
            bit >>= 1;
        }
        float v = s[0];
        for (; bit > 0; bit >>= 1) v += __shfl_xor_sync(0xffffffffu, v, bit);
        const int owner_mask = TQ == 2 ? 3 : 7;                                  // lanes that differ only in the low bits hold the same sum
        if ((lane & owner_mask) == 0) {
            const int u = idx / TQ, i = idx % TQ;
            if (r0 + 8 * u < r_lim) out[i * out_stride + r0 + 8 * u] = v;
        }
    };
    for (int r0 = warp; r0 < r_end; r0 += 32) dots4(Ps, r0, r_end, qv, &sf.bd[0][0], ATT_MAX_REL);
    for (int j0 = first + warp; j0 < K; j0 += 32) dots4(Ks, j0, K, qu, &sf.ac[0][0], ATT_MAX_K);
    half_barrier(half);
    // ---- softmax over valid keys; rel-shift = index arithmetic: BD[i][j] = BD_raw[i][L + i - j + T-1] (:391-433) ----
    for (int i = warp; i < T; i += 8) {
        float mx = -INFINITY;
        for (int j = first + lane; j < K; j += 32) {
            const float s = (sf.ac[i][j] + sf.bd[i][ATT_L + i - j + T - 1]) * scale;
            sf.ac[i][j] = s; mx = fmaxf(mx, s);
        }
        mx = warp_max(mx);
        float sum = 0.f;
        for (int j = first + lane; j < K; j += 32) { const float e = expf(sf.ac[i][j] - mx); sf.ac[i][j] = e; sum += e; }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
        for (int j = first + lane; j < K; j += 32) sf.ac[i][j] *= inv;
    }
    half_barrier(half);
    if (tr_slot >= 0) trace_mark(tr_slot, 4);
    // ---- ctx[i][d] = sum_j p[i][j] v_j[d]; thread = (key group kg, 2 dims) ----
    {
        const int d2 = (t & 63) * 2, kg = t >> 6;
        float acc[TQ][2];
#pragma unroll
        for (int i = 0; i < TQ; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; }
#pragma unroll 4
        for (int j = first + kg; j < K; j += 4) {
            float vf[2];
            load2(Vs + (size_t)j * D_HEAD + d2, vf);
#pragma unroll
            for (int i = 0; i < TQ; ++i) {
                const float p = sf.ac[i][j];
                acc[i][0] = fmaf(p, vf[0], acc[i][0]); acc[i][1] = fmaf(p, vf[1], acc[i][1]);
            }
        }
#pragma unroll
        for (int i = 0; i < TQ; ++i) { sf.red[kg][i][d2] = acc[i][0]; sf.red[kg][i][d2 + 1] = acc[i][1]; }
    }
    half_barrier(half);
    for (int e = t; e < T * D_HEAD; e += 256) {
        const int i = e / D_HEAD, d = e % D_HEAD;
        const float v = (sf.red[0][i][d] + sf.red[1][i][d]) + (sf.red[2][i][d] + sf.red[3][i][d]);
        store_out(a.ctx, ((size_t)b * T + i) * D_MODEL + h * D_HEAD + d, v, a.out_type);
    }
    NSB_KERNEL_EPILOGUE();
}

template <int KV, int TQ>
static void launch_attention_pair(const AttnArgs& a, cudaStream_t st) {
    using E = typename KvT<KV>::type;
    const size_t smem = (size_t)(ATT_L + 2 * a.T - 1) * D_HEAD * sizeof(E) + 2 * (sizeof(AttnSmemF<TQ>) + (size_t)2 * (ATT_L + a.T) * D_HEAD * sizeof(E));
    static std::atomic<size_t> configured[MAX_DEVICES];
    ensure_dyn_smem(attention_pair_kernel<KV, TQ>, smem, configured);
    launch_k(attention_pair_kernel<KV, TQ>, dim3(N_HEADS, (a.B + 1) / 2), dim3(512), smem, st, a);
}
static bool attention_pair_enabled() {
    static const bool on = [] { const char* e = getenv("NSB_ATT_PAIR"); return !(e && e[0] == '0'); }();
    return on;
}

// ------------------------------------------------------------------------------------------
// Tensor-core attention for a 16-bit ring and T <= 16 (all four latency modes). The three small GEMMs of a (head, stream) --
//   AC = (q+u) K^T  [T x K],   BD_raw = (q+v) P^T  [T x (L+2T-1)],   ctx = softmax(..) V  [T x 128]
// -- run as mma.sync.m16n8k16 (the T query rows sit in one 16-row tile, rows >= T read a shared zero row) with fp32
// accumulation; operands come out of shared memory with ldmatrix (rows padded to 272 bytes: conflict-free). The scalar kernels
// above spend ~600 instructions per warp on per-lane partial dot products and warp reductions (issue bound; at T = 7 the
// 2048 CTAs of a 256-stream step take 90 us per layer); here a stream needs ~250 HMMA in total.
// (q+u), (q+v) and the probabilities are rounded to the ring dtype before they enter the tensor core -- the same precision
// class as K, V and P themselves. One CTA = one head x NS streams (NS = 2 for T <= 2: the positional rows are staged once per
// CTA), 4 warps per stream; everything that does not depend on this step's QKV GEMM is fetched before the dependency wait.
// ------------------------------------------------------------------------------------------
constexpr int ATT_RS = D_HEAD + 8;                                // padded row stride (elements) of the 16-bit tiles: 272 bytes

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t (&r)[2], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];"
                 : "=r"(r[0]), "=r"(r[1]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
template <typename E> __device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1);
template <> __device__ __forceinline__ void mma_16816<__half>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <> __device__ __forceinline__ void mma_16816<__nv_bfloat16>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void group_barrier(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

struct AttnMmaLayout {                                            // byte offsets inside the dynamic shared memory
    int kpad, rpad, n_rel;                                        // keys padded to 16, positional rows padded to 8
    size_t ps, zero, per_stream, ks, vs, qu, qv, pat, ac, bd, total;
    // dv (single-stream CTAs): the V tile is loaded AFTER the two score GEMMs into the space of the positional tile, which is dead
    // by then -- 54 KB instead of 77 KB per CTA = 4 resident CTAs per SM instead of 2 (vs is then an absolute offset)
    __host__ __device__ AttnMmaLayout(int T, int ns, bool dv) {
        const int K = ATT_L + T;
        n_rel = ATT_L + 2 * T - 1; kpad = (K + 15) & ~15; rpad = (n_rel + 7) & ~7;
        const size_t row = ATT_RS * 2;
        ps = 0; zero = ps + (size_t)(rpad > kpad ? rpad : kpad) * row;
        const size_t base = zero + row;
        ks = 0; vs = ks + (size_t)kpad * row; qu = dv ? vs : vs + (size_t)kpad * row; qv = qu + (size_t)T * row; pat = qv + (size_t)T * row;
        ac = pat + (size_t)T * (kpad + 8) * 2; ac = (ac + 15) & ~(size_t)15;
        bd = ac + (size_t)T * kpad * 4; per_stream = bd + (size_t)T * rpad * 4; per_stream = (per_stream + 15) & ~(size_t)15;
        ks += base; vs = dv ? ps : vs + base; qu += base; qv += base; pat += base; ac += base; bd += base;
        total = base + (size_t)ns * per_stream;
    }
};

template <int KV, int NS>
__global__ void __launch_bounds__(128 * NS, NS == 2 ? 2 : 4) attention_mma_kernel(const AttnArgs a) {
    constexpr bool DV = NS == 1;                                                // deferred V tile (see AttnMmaLayout)
    using E = typename KvT<KV>::type;
    static_assert(KV != 0, "16-bit ring only");
    constexpr int EPV = 8, VPR = D_HEAD / EPV;                                  // 16-byte vectors: 8 elements, 16 per row
    extern __shared__ __align__(16) uint8_t att_smem[];
    const int T = a.T, K = ATT_L + T, Cap = K;
    const AttnMmaLayout lay(T, NS, DV);
    const int tid = threadIdx.x, half = NS == 2 ? tid >> 7 : 0, t = tid & 127, warp = t >> 5, lane = t & 31;
    const int h = blockIdx.x, b = blockIdx.y * NS + half;
    const bool active = b < a.B;
    uint8_t* hb = att_smem + (size_t)half * lay.per_stream;
    E* Ps = reinterpret_cast<E*>(att_smem + lay.ps);                            // [rpad][ATT_RS], shared by the streams of the CTA
    E* Zero = reinterpret_cast<E*>(att_smem + lay.zero);                        // one all-zero row
    E* Ks = reinterpret_cast<E*>(hb + lay.ks);                                  // [kpad][ATT_RS]
    E* Vs = reinterpret_cast<E*>((DV ? att_smem : hb) + lay.vs);
    E* Qu = reinterpret_cast<E*>(hb + lay.qu);                                  // [T][ATT_RS]  (q + pos_bias_u), rounded
    E* Qv = reinterpret_cast<E*>(hb + lay.qv);
    E* Pat = reinterpret_cast<E*>(hb + lay.pat);                                // [T][kpad + 8] probabilities, rounded
    float* Ac = reinterpret_cast<float*>(hb + lay.ac);                          // [T][kpad]
    float* Bd = reinterpret_cast<float*>(hb + lay.bd);                          // [T][rpad]
    const int kpad = lay.kpad, rpad = lay.rpad, n_rel = lay.n_rel, pat_rs = kpad + 8;
    NSB_KERNEL_BEGIN(TR_ATTN)
    // ---- before the dependency wait: positional rows, cached K / V rows, zero fills ----
    const E* P = reinterpret_cast<const E*>(a.pos_proj) + h * D_HEAD;
    for (int e = tid; e < rpad * VPR; e += 128 * NS) {
        const int r = e / VPR, c = (e % VPR) * EPV;
        if (r < n_rel) cp_async16(Ps + (size_t)r * ATT_RS + c, P + (size_t)r * D_MODEL + c);
        else *reinterpret_cast<uint4*>(Ps + (size_t)r * ATT_RS + c) = make_uint4(0, 0, 0, 0);
    }
    for (int e = tid; e < VPR; e += 128 * NS) *reinterpret_cast<uint4*>(Zero + e * EPV) = make_uint4(0, 0, 0, 0);
    int slot = 0, w = 0, first = ATT_L;
    if (active) { slot = a.slot_of_b[b]; w = a.ring_pos[slot]; first = ATT_L - a.valid_len[slot]; }   // keys j < first are not yet valid (:982-992)
    const float* qkv = a.qkv + (size_t)b * T * 3 * D_MODEL + h * D_HEAD;
    E* kring = reinterpret_cast<E*>(a.k_ring) + (size_t)slot * a.slot_stride + h * D_HEAD;
    E* vring = reinterpret_cast<E*>(a.v_ring) + (size_t)slot * a.slot_stride + h * D_HEAD;
    const int rbase = w + Cap - ATT_L;                                            // ring row of key j: (rbase + j) mod Cap, without a division
    auto ring_row = [&](int j) { int r = rbase + j; r -= r >= Cap ? Cap : 0; r -= r >= Cap ? Cap : 0; return (size_t)r * D_MODEL; };
    if (active) {
        for (int e = t; e < kpad * VPR; e += 128) {
            const int j = e / VPR, c = (e % VPR) * EPV;
            if (j >= first && j < ATT_L) {
                const size_t g = ring_row(j) + c;
                cp_async16(Ks + (size_t)j * ATT_RS + c, kring + g);
                if (!DV) cp_async16(Vs + (size_t)j * ATT_RS + c, vring + g);
                else if ((c & 63) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(vring + g));   // V rows: HBM -> L2 now, shared memory after the score GEMMs
            } else if (j < first || j >= K) {                                    // not yet valid / padding: zeros (0 x garbage must not be NaN)
                *reinterpret_cast<uint4*>(Ks + (size_t)j * ATT_RS + c) = make_uint4(0, 0, 0, 0);
                if (!DV) *reinterpret_cast<uint4*>(Vs + (size_t)j * ATT_RS + c) = make_uint4(0, 0, 0, 0);
            }
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    // q / k / v staging: a thread owns two adjacent head dims (t2) of every second row (rsel)
    struct alignas(4) E2 { E a, b; };
    const int t2 = (t & 63) * 2, rsel = t >> 6;
    const float2 bu = *reinterpret_cast<const float2*>(a.bias_u + h * D_HEAD + t2), bv = *reinterpret_cast<const float2*>(a.bias_v + h * D_HEAD + t2);
    NSB_KERNEL_WAIT()
    if (active) {
        for (int i0 = 0; i0 < T; i0 += 8) {                                      // 4 rows per thread = 12 loads in flight
            float2 q[4], kn[4], vn[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + rsel + 2 * u;
                if (i >= T) continue;
                const float* row = qkv + (size_t)i * 3 * D_MODEL + t2;
                q[u] = *reinterpret_cast<const float2*>(row); kn[u] = *reinterpret_cast<const float2*>(row + D_MODEL); vn[u] = *reinterpret_cast<const float2*>(row + 2 * D_MODEL);
                for (int z = 1; z < a.planes; ++z) {                             // split-K partial planes of the QKV GEMM, in slice order
                    const float* rz = row + (size_t)z * a.plane_stride;
                    const float2 q2 = *reinterpret_cast<const float2*>(rz), k2 = *reinterpret_cast<const float2*>(rz + D_MODEL), v2 = *reinterpret_cast<const float2*>(rz + 2 * D_MODEL);
                    q[u].x += q2.x; q[u].y += q2.y; kn[u].x += k2.x; kn[u].y += k2.y; vn[u].x += v2.x; vn[u].y += v2.y;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + rsel + 2 * u;
                if (i >= T) continue;
                *reinterpret_cast<E2*>(Qu + (size_t)i * ATT_RS + t2) = E2{from_f32<E>(q[u].x + bu.x), from_f32<E>(q[u].y + bu.y)};   // :503-507
                *reinterpret_cast<E2*>(Qv + (size_t)i * ATT_RS + t2) = E2{from_f32<E>(q[u].x + bv.x), from_f32<E>(q[u].y + bv.y)};
                const E2 ke{from_f32<E>(kn[u].x), from_f32<E>(kn[u].y)}, ve{from_f32<E>(vn[u].x), from_f32<E>(vn[u].y)};
                *reinterpret_cast<E2*>(Ks + (size_t)(ATT_L + i) * ATT_RS + t2) = ke;
                if (!DV) *reinterpret_cast<E2*>(Vs + (size_t)(ATT_L + i) * ATT_RS + t2) = ve;
                const size_t g = ring_row(ATT_L + i) + t2;                       // ring append (replaces concat + roll, :465-484)
                *reinterpret_cast<E2*>(kring + g) = ke; *reinterpret_cast<E2*>(vring + g) = ve;
            }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                                              // the only CTA-wide barrier: P / K / V tiles, q and the new rows are in place
    if (tr_slot >= 0) trace_mark(tr_slot, 3);
    if (!active) return;
    const int g = lane >> 2, tq = lane & 3;                                       // mma fragment coordinates
    // per-lane ldmatrix row addresses of the A operand (16 x 16 tile): row = lane % 8 + 8 * (lane / 8 % 2), column block = lane / 16
    const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8, a_col = (lane >> 4) * 8;
    // ---- AC = Qu K^T and BD_raw = Qv P^T: 8-wide column tiles round-robin over the 4 warps ----
    {
        const int nt_ac = (K + 7) / 8, nt_bd = rpad / 8;
        const E* qa_u = a_row < T ? Qu + (size_t)a_row * ATT_RS + a_col : Zero + a_col;
        const E* qa_v = a_row < T ? Qv + (size_t)a_row * ATT_RS + a_col : Zero + a_col;
        // two column tiles per warp step: two independent accumulator chains hide the ldmatrix / mma latency
        const int ntot = nt_ac + nt_bd;
        for (int tile0 = warp; tile0 < ntot; tile0 += 8) {
            float c[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
            const E* bp[2]; const E* qa[2]; bool live[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int tile = tile0 + 4 * u;
                live[u] = tile < ntot;
                const bool is_ac = tile < nt_ac;
                const int n0 = live[u] ? (is_ac ? tile : tile - nt_ac) * 8 : 0;
                // B operand: lanes 0-7 -> rows n0.., k 0-7; 8-15 -> k 8-15; 16-23 -> k 16-23; 24-31 -> k 24-31 (two k-steps per ldmatrix.x4)
                bp[u] = (is_ac ? Ks : Ps) + (size_t)(n0 + (lane & 7)) * ATT_RS + (lane >> 3) * 8;
                qa[u] = is_ac ? qa_u : qa_v;
            }
#pragma unroll
            for (int ks = 0; ks < D_HEAD / 32; ++ks) {
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    uint32_t bf[4], a0[4], a1[4];
                    ldsm_x4(bf, bp[u] + ks * 32);
                    ldsm_x4(a0, qa[u] + ks * 32);
                    ldsm_x4(a1, qa[u] + ks * 32 + 16);
                    mma_16816<E>(c[u], a0, bf[0], bf[1]);
                    mma_16816<E>(c[u], a1, bf[2], bf[3]);
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int tile = tile0 + 4 * u;
                if (!live[u]) continue;
                const bool is_ac = tile < nt_ac;
                const int n0 = (is_ac ? tile : tile - nt_ac) * 8;
                float* out = is_ac ? Ac : Bd;
                const int ld = is_ac ? kpad : rpad;
                if (g < T) { out[g * ld + n0 + 2 * tq] = c[u][0]; out[g * ld + n0 + 2 * tq + 1] = c[u][1]; }
                if (g + 8 < T) { out[(g + 8) * ld + n0 + 2 * tq] = c[u][2]; out[(g + 8) * ld + n0 + 2 * tq + 1] = c[u][3]; }
            }
        }
    }
    group_barrier(1 + half, 128);
    if (DV) {                                                                     // the positional tile is dead: V (cached rows + this chunk's rows, now in the ring) takes its place
        for (int e = t; e < kpad * VPR; e += 128) {
            const int j = e / VPR, c = (e % VPR) * EPV;
            if (j >= first && j < K) cp_async16(Vs + (size_t)j * ATT_RS + c, vring + ring_row(j) + c);
            else *reinterpret_cast<uint4*>(Vs + (size_t)j * ATT_RS + c) = make_uint4(0, 0, 0, 0);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    // ---- softmax over valid keys; rel-shift = index arithmetic: BD[i][j] = BD_raw[i][L + i - j + T-1] (:391-433) ----
    const float scale = 0.08838834764831845f;                                   // 1/sqrt(128) :517
    for (int i = warp; i < T; i += 4) {
        float mx = -INFINITY;
        for (int j = first + lane; j < K; j += 32) {
            const float sc = (Ac[i * kpad + j] + Bd[i * rpad + ATT_L + i - j + T - 1]) * scale;
            Ac[i * kpad + j] = sc; mx = fmaxf(mx, sc);
        }
        mx = warp_max(mx);
        float sum = 0.f;
        for (int j = first + lane; j < K; j += 32) { const float e = expf(Ac[i * kpad + j] - mx); Ac[i * kpad + j] = e; sum += e; }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
        for (int j = lane; j < kpad; j += 32)
            Pat[(size_t)i * pat_rs + j] = from_f32<E>(j >= first && j < K ? Ac[i * kpad + j] * inv : 0.f);
    }
    if (DV) asm volatile("cp.async.wait_group 0;" ::: "memory");
    group_barrier(1 + half, 128);
    if (tr_slot >= 0) trace_mark(tr_slot, 4);
    // ---- ctx = P V: 16 output tiles of 8 head dims, 4 per warp; k runs over the (padded) keys ----
    {
        const E* pa = a_row < T ? Pat + (size_t)a_row * pat_rs + a_col : Zero + a_col;   // zero row is 136 elements: a_col + 16 ks stays inside for kpad <= 96
        // the warp's four output tiles (head dims 8 (warp + 4 u) ..) together: the probability fragment is loaded once per k-step
        float c[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { c[u][0] = 0.f; c[u][1] = 0.f; c[u][2] = 0.f; c[u][3] = 0.f; }
        for (int ks = 0; ks < kpad / 16; ++ks) {
            uint32_t af[4];
            ldsm_x4(af, pa + ks * 16);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                uint32_t bf[2];
                ldsm_x2_trans(bf, Vs + (size_t)(ks * 16 + (lane & 15)) * ATT_RS + (warp + 4 * u) * 8);   // lanes 0-7: keys k0..k0+7, lanes 8-15: k0+8..k0+15
                mma_16816<E>(c[u], af, bf[0], bf[1]);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int n0 = (warp + 4 * u) * 8;
            if (g < T) {
                const size_t o = ((size_t)b * T + g) * D_MODEL + h * D_HEAD + n0 + 2 * tq;
                store_out(a.ctx, o, c[u][0], a.out_type); store_out(a.ctx, o + 1, c[u][1], a.out_type);
            }
            if (g + 8 < T) {
                const size_t o = ((size_t)b * T + g + 8) * D_MODEL + h * D_HEAD + n0 + 2 * tq;
                store_out(a.ctx, o, c[u][2], a.out_type); store_out(a.ctx, o + 1, c[u][3], a.out_type);
            }
        }
    }
    NSB_KERNEL_EPILOGUE();
}

template <int KV, int NS>
static void launch_attention_mma(const AttnArgs& a, cudaStream_t st) {
    const AttnMmaLayout lay(a.T, NS, NS == 1);
    static std::atomic<size_t> configured[MAX_DEVICES];
    // largest shared-memory carve-out: the resident CTAs per SM are limited by shared memory, not by L1
    ensure_dyn_smem(attention_mma_kernel<KV, NS>, lay.total, configured, true);
    launch_k(attention_mma_kernel<KV, NS>, dim3(N_HEADS, (a.B + NS - 1) / NS), dim3(128 * NS), lay.total, st, a);
}

// ------------------------------------------------------------------------------------------
// The same arithmetic for LARGE batches (8 x B >> resident CTAs; 256 streams x 7 frames = 2048 (head, stream) items): the kernel above
// hides its load -> score -> load -> softmax -> PV chain only through occupancy (4 CTAs per SM, 3.5 waves, 42 us per layer for the
// 13 us of ring bytes). Here a CTA keeps ONE head, stages that head's positional rows once and walks its streams with the K / V tiles
// of the next stream (cp.async into a three-tile ring: K_i, V_i, K_i+1; V_i+1 takes K_i's place after the scores) and the next
// stream's q / k / v rows (registers) in flight under the current stream's MMAs. Two CTAs per SM. Same instructions per item in the
// same order as attention_mma_kernel<KV, 1>: bit-identical output.
// ------------------------------------------------------------------------------------------
struct AttnStreamLayout {
    int kpad, rpad, n_rel;
    size_t ps, zero, tile0, tile_bytes, qu, qv, pat, ac, bd, total;          // three K / V tiles from tile0
    __host__ __device__ AttnStreamLayout(int T) {
        const int K = ATT_L + T;
        n_rel = ATT_L + 2 * T - 1; kpad = (K + 15) & ~15; rpad = (n_rel + 7) & ~7;
        const size_t row = ATT_RS * 2;
        ps = 0; zero = ps + (size_t)rpad * row;
        size_t o = zero + row;
        tile0 = o; tile_bytes = (size_t)kpad * row; o += 3 * tile_bytes;
        qu = o; o += (size_t)T * row; qv = o; o += (size_t)T * row;
        pat = o; o += (size_t)T * (kpad + 8) * 2; o = (o + 15) & ~(size_t)15;
        ac = o; o += (size_t)T * kpad * 4; bd = o; o += (size_t)T * rpad * 4;
        total = (o + 15) & ~(size_t)15;
    }
};
constexpr int ATT_STREAM_MAX_T = 8;                                               // one pass of the 4-rows-per-thread q / k / v staging

// this chunk's q / k / v rows of one (stream, head): fp32 GEMM output, split-K planes folded in slice order; a thread owns two adjacent
// head dims (t2) of every second row (rsel)
__device__ __forceinline__ void att_load_rows(const AttnArgs& a, int bb, int h, int T, int rsel, int t2, float2 (&q)[4], float2 (&kn)[4], float2 (&vn)[4]) {
    const float* qkv = a.qkv + (size_t)bb * T * 3 * D_MODEL + h * D_HEAD;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int i = rsel + 2 * u;
        if (i < T) {
            const float* row = qkv + (size_t)i * 3 * D_MODEL + t2;
            q[u] = *reinterpret_cast<const float2*>(row); kn[u] = *reinterpret_cast<const float2*>(row + D_MODEL); vn[u] = *reinterpret_cast<const float2*>(row + 2 * D_MODEL);
            for (int z = 1; z < a.planes; ++z) {
                const float* rz = row + (size_t)z * a.plane_stride;
                const float2 q2 = *reinterpret_cast<const float2*>(rz), k2 = *reinterpret_cast<const float2*>(rz + D_MODEL), v2 = *reinterpret_cast<const float2*>(rz + 2 * D_MODEL);
                q[u].x += q2.x; q[u].y += q2.y; kn[u].x += k2.x; kn[u].y += k2.y; vn[u].x += v2.x; vn[u].y += v2.y;
            }
        }
    }
}

template <int KV>
__global__ void __launch_bounds__(128, 2) attention_mma_stream_kernel(const AttnArgs a, int G) {
    using E = typename KvT<KV>::type;
    static_assert(KV != 0, "16-bit ring only");
    constexpr int EPV = 8, VPR = D_HEAD / EPV;
    extern __shared__ __align__(16) uint8_t att_smem[];
    const int T = a.T, K = ATT_L + T, Cap = K, B = a.B;
    const AttnStreamLayout lay(T);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int h = blockIdx.x;
    E* Ps = reinterpret_cast<E*>(att_smem + lay.ps);
    E* Zero = reinterpret_cast<E*>(att_smem + lay.zero);
    E* Qu = reinterpret_cast<E*>(att_smem + lay.qu);
    E* Qv = reinterpret_cast<E*>(att_smem + lay.qv);
    E* Pat = reinterpret_cast<E*>(att_smem + lay.pat);
    float* Ac = reinterpret_cast<float*>(att_smem + lay.ac);
    float* Bd = reinterpret_cast<float*>(att_smem + lay.bd);
    const int kpad = lay.kpad, rpad = lay.rpad, n_rel = lay.n_rel, pat_rs = kpad + 8;
    NSB_KERNEL_BEGIN(TR_ATTN)
    const E* P = reinterpret_cast<const E*>(a.pos_proj) + h * D_HEAD;
    for (int e = tid; e < rpad * VPR; e += 128) {
        const int r = e / VPR, c = (e % VPR) * EPV;
        if (r < n_rel) cp_async16(Ps + (size_t)r * ATT_RS + c, P + (size_t)r * D_MODEL + c);
        else *reinterpret_cast<uint4*>(Ps + (size_t)r * ATT_RS + c) = make_uint4(0, 0, 0, 0);
    }
    for (int e = tid; e < VPR; e += 128) *reinterpret_cast<uint4*>(Zero + e * EPV) = make_uint4(0, 0, 0, 0);
    struct St { int slot, rbase, first; };                                        // per-stream ring state
    auto load_st = [&](int b) {
        St s; s.slot = a.slot_of_b[b];
        s.rbase = a.ring_pos[s.slot] + Cap - ATT_L;                               // ring row of key j: (rbase + j) mod Cap
        s.first = ATT_L - a.valid_len[s.slot];                                    // keys j < first are not yet valid (:982-992)
        return s;
    };
    auto ring_row = [&](const St& s, int j) { int r = s.rbase + j; r -= r >= Cap ? Cap : 0; r -= r >= Cap ? Cap : 0; return (size_t)r * D_MODEL; };
    // cached rows of one tile: cp.async for the valid ones, zeros for the not-yet-valid and the padding rows (rows ATT_L .. K-1 = this
    // chunk's rows are written by the staging step)
    auto issue_tile = [&](E* dst, void* ring_base, const St& s) {
        const E* ring = reinterpret_cast<const E*>(ring_base) + (size_t)s.slot * a.slot_stride + h * D_HEAD;
        for (int e = tid; e < kpad * VPR; e += 128) {
            const int j = e / VPR, c = (e % VPR) * EPV;
            if (j >= s.first && j < ATT_L) cp_async16(dst + (size_t)j * ATT_RS + c, ring + ring_row(s, j) + c);
            else if (j < s.first || j >= K) *reinterpret_cast<uint4*>(dst + (size_t)j * ATT_RS + c) = make_uint4(0, 0, 0, 0);
        }
    };
    int b = blockIdx.y, kt = 0, vt = 1;
    St cur = load_st(b), nxt = cur;
    issue_tile(reinterpret_cast<E*>(att_smem + lay.tile0), a.k_ring, cur);
    asm volatile("cp.async.commit_group;" ::: "memory");                        // group: P + K_0
    issue_tile(reinterpret_cast<E*>(att_smem + lay.tile0 + lay.tile_bytes), a.v_ring, cur);
    asm volatile("cp.async.commit_group;" ::: "memory");                        // group: V_0
    if (b + G < B) nxt = load_st(b + G);
    struct alignas(4) E2 { E a, b; };
    const int t2 = (tid & 63) * 2, rsel = tid >> 6;
    const float2 bu = *reinterpret_cast<const float2*>(a.bias_u + h * D_HEAD + t2), bv = *reinterpret_cast<const float2*>(a.bias_v + h * D_HEAD + t2);
    float2 q[4], kn[4], vn[4];
    NSB_KERNEL_WAIT()
    att_load_rows(a, b, h, T, rsel, t2, q, kn, vn);
    const int g = lane >> 2, tq = lane & 3;
    const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8, a_col = (lane >> 4) * 8;
    const float scale = 0.08838834764831845f;                                   // 1/sqrt(128) :517
    bool first_item = true;
    for (; b < B; b += G) {
        E* Ks = reinterpret_cast<E*>(att_smem + lay.tile0 + kt * lay.tile_bytes);
        E* Vs = reinterpret_cast<E*>(att_smem + lay.tile0 + vt * lay.tile_bytes);
        const bool has_next = b + G < B;
        // ---- 1. stage this chunk's rows: (q + u), (q + v), new K / V rows (shared memory + ring append :465-484) ----
        {
            E* kring = reinterpret_cast<E*>(a.k_ring) + (size_t)cur.slot * a.slot_stride + h * D_HEAD;
            E* vring = reinterpret_cast<E*>(a.v_ring) + (size_t)cur.slot * a.slot_stride + h * D_HEAD;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = rsel + 2 * u;
                if (i < T) {
                    *reinterpret_cast<E2*>(Qu + (size_t)i * ATT_RS + t2) = E2{from_f32<E>(q[u].x + bu.x), from_f32<E>(q[u].y + bu.y)};   // :503-507
                    *reinterpret_cast<E2*>(Qv + (size_t)i * ATT_RS + t2) = E2{from_f32<E>(q[u].x + bv.x), from_f32<E>(q[u].y + bv.y)};
                    const E2 ke{from_f32<E>(kn[u].x), from_f32<E>(kn[u].y)}, ve{from_f32<E>(vn[u].x), from_f32<E>(vn[u].y)};
                    *reinterpret_cast<E2*>(Ks + (size_t)(ATT_L + i) * ATT_RS + t2) = ke;
                    *reinterpret_cast<E2*>(Vs + (size_t)(ATT_L + i) * ATT_RS + t2) = ve;
                    const size_t gr = ring_row(cur, ATT_L + i) + t2;
                    *reinterpret_cast<E2*>(kring + gr) = ke; *reinterpret_cast<E2*>(vring + gr) = ve;
                }
            }
        }
        // ---- 2. next stream: K tile into the buffer V_{i-1} left, q / k / v rows into registers ----
        if (has_next) { issue_tile(reinterpret_cast<E*>(att_smem + lay.tile0 + ((kt + 2) % 3) * lay.tile_bytes), a.k_ring, nxt); att_load_rows(a, b + G, h, T, rsel, t2, q, kn, vn); }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 2;" ::: "memory");                     // P and K_i have landed (V_i, K_i+1 may be in flight)
        __syncthreads();
        if (first_item && tr_slot >= 0) trace_mark(tr_slot, 3);
        // ---- 3. AC = Qu K^T and BD_raw = Qv P^T ----
        {
            const int nt_ac = (K + 7) / 8, nt_bd = rpad / 8;
            const E* qa_u = a_row < T ? Qu + (size_t)a_row * ATT_RS + a_col : Zero + a_col;
            const E* qa_v = a_row < T ? Qv + (size_t)a_row * ATT_RS + a_col : Zero + a_col;
            const int ntot = nt_ac + nt_bd;
            for (int tile0 = warp; tile0 < ntot; tile0 += 8) {
                float c[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
                const E* bp[2]; const E* qa[2]; bool live[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int tile = tile0 + 4 * u;
                    live[u] = tile < ntot;
                    const bool is_ac = tile < nt_ac;
                    const int n0 = live[u] ? (is_ac ? tile : tile - nt_ac) * 8 : 0;
                    bp[u] = (is_ac ? Ks : Ps) + (size_t)(n0 + (lane & 7)) * ATT_RS + (lane >> 3) * 8;
                    qa[u] = is_ac ? qa_u : qa_v;
                }
#pragma unroll
                for (int ks = 0; ks < D_HEAD / 32; ++ks) {
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        uint32_t bf[4], a0[4], a1[4];
                        ldsm_x4(bf, bp[u] + ks * 32);
                        ldsm_x4(a0, qa[u] + ks * 32);
                        ldsm_x4(a1, qa[u] + ks * 32 + 16);
                        mma_16816<E>(c[u], a0, bf[0], bf[1]);
                        mma_16816<E>(c[u], a1, bf[2], bf[3]);
                    }
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int tile = tile0 + 4 * u;
                    if (!live[u]) continue;
                    const bool is_ac = tile < nt_ac;
                    const int n0 = (is_ac ? tile : tile - nt_ac) * 8;
                    float* out = is_ac ? Ac : Bd;
                    const int ld = is_ac ? kpad : rpad;
                    if (g < T) { out[g * ld + n0 + 2 * tq] = c[u][0]; out[g * ld + n0 + 2 * tq + 1] = c[u][1]; }
                    if (g + 8 < T) { out[(g + 8) * ld + n0 + 2 * tq] = c[u][2]; out[(g + 8) * ld + n0 + 2 * tq + 1] = c[u][3]; }
                }
            }
        }
        __syncthreads();                                                          // K_i is dead
        // ---- 4. next stream's V tile into K_i's buffer ----
        if (has_next) issue_tile(Ks, a.v_ring, nxt);
        asm volatile("cp.async.commit_group;" ::: "memory");
        // ---- 5. softmax over valid keys; rel-shift = index arithmetic: BD[i][j] = BD_raw[i][L + i - j + T-1] (:391-433) ----
        const int first = cur.first;
        for (int i = warp; i < T; i += 4) {
            float mx = -INFINITY;
            for (int j = first + lane; j < K; j += 32) {
                const float sc = (Ac[i * kpad + j] + Bd[i * rpad + ATT_L + i - j + T - 1]) * scale;
                Ac[i * kpad + j] = sc; mx = fmaxf(mx, sc);
            }
            mx = warp_max(mx);
            float sum = 0.f;
            for (int j = first + lane; j < K; j += 32) { const float e = expf(Ac[i * kpad + j] - mx); Ac[i * kpad + j] = e; sum += e; }
            sum = warp_sum(sum);
            const float inv = 1.0f / sum;
            for (int j = lane; j < kpad; j += 32)
                Pat[(size_t)i * pat_rs + j] = from_f32<E>(j >= first && j < K ? Ac[i * kpad + j] * inv : 0.f);
        }
        asm volatile("cp.async.wait_group 2;" ::: "memory");                     // V_i has landed (K_i+1, V_i+1 may be in flight)
        __syncthreads();
        if (first_item && tr_slot >= 0) trace_mark(tr_slot, 4);
        // ---- 6. ctx = P V ----
        {
            const E* pa = a_row < T ? Pat + (size_t)a_row * pat_rs + a_col : Zero + a_col;
            float c[4][4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { c[u][0] = 0.f; c[u][1] = 0.f; c[u][2] = 0.f; c[u][3] = 0.f; }
            for (int ks = 0; ks < kpad / 16; ++ks) {
                uint32_t af[4];
                ldsm_x4(af, pa + ks * 16);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    uint32_t bf[2];
                    ldsm_x2_trans(bf, Vs + (size_t)(ks * 16 + (lane & 15)) * ATT_RS + (warp + 4 * u) * 8);
                    mma_16816<E>(c[u], af, bf[0], bf[1]);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int n0 = (warp + 4 * u) * 8;
                if (g < T) {
                    const size_t o = ((size_t)b * T + g) * D_MODEL + h * D_HEAD + n0 + 2 * tq;
                    store_out(a.ctx, o, c[u][0], a.out_type); store_out(a.ctx, o + 1, c[u][1], a.out_type);
                }
                if (g + 8 < T) {
                    const size_t o = ((size_t)b * T + g + 8) * D_MODEL + h * D_HEAD + n0 + 2 * tq;
                    store_out(a.ctx, o, c[u][2], a.out_type); store_out(a.ctx, o + 1, c[u][3], a.out_type);
                }
            }
        }
        __syncthreads();                                                          // V_i, Qu / Qv / Pat / Ac / Bd are free for the next stream
        cur = nxt;
        if (b + 2 * G < B) nxt = load_st(b + 2 * G);
        const int k_next = (kt + 2) % 3; vt = kt; kt = k_next;
        first_item = false;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    NSB_KERNEL_EPILOGUE();
}

template <int KV>
static void launch_attention_mma_stream(const AttnArgs& a, cudaStream_t st) {
    const AttnStreamLayout lay(a.T);
    static std::atomic<size_t> configured[MAX_DEVICES];
    ensure_dyn_smem(attention_mma_stream_kernel<KV>, lay.total, configured, true);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int slots = std::max(1, 2 * sms / N_HEADS);                             // CTA columns that are resident at once (two CTAs per SM)
    const int per = (a.B + slots - 1) / slots, G = (a.B + per - 1) / per;         // streams per CTA, evened out
    launch_k(attention_mma_stream_kernel<KV>, dim3(N_HEADS, G), dim3(128), lay.total, st, a, G);
}
// NSB_ATT_STREAM=1: whenever T fits. EXPERIMENT, off by default: measured SLOWER than the one-item kernel at 256 streams x 7 frames
// (62 vs 51 us per launch): the item is not bound by its load latency.
static bool attention_stream_wanted(const AttnArgs& a) {
    if (a.T < 3 || a.T > ATT_STREAM_MAX_T) return false;
    const char* e = getenv("NSB_ATT_STREAM");
    return e && e[0] == '1';
}
static bool attention_mma_enabled() {
    static const bool on = [] { const char* e = getenv("NSB_ATT_MMA"); return !(e && e[0] == '0'); }();
    return on;
}

template <int KV, int TQ>
static void launch_attention_tq(const AttnArgs& a, cudaStream_t st) {
    using E = typename KvT<KV>::type;
    const size_t smem = sizeof(AttnSmemF<TQ>) + (size_t)2 * (ATT_L + a.T) * D_HEAD * sizeof(E);
    static std::atomic<size_t> configured[MAX_DEVICES];
    ensure_dyn_smem(attention_kernel<KV, TQ>, smem, configured);
    launch_k(attention_kernel<KV, TQ>, dim3(N_HEADS, a.B), dim3(256), smem, st, a);
}
template <int KV>
static void launch_attention_t(const AttnArgs& a, cudaStream_t st) {
    if (a.T > ATT_MAX_T) throw CudaError("attention: att_right_context too large");
    if constexpr (KV != 0) {                                                      // 16-bit ring: tensor-core kernel (T <= 16), else the paired / tiled scalar kernels
        if (a.T <= 16 && attention_mma_enabled()) {
            static const bool ns1 = [] { const char* e = getenv("NSB_ATT_NS1"); return e && e[0] == '1'; }();   // experiment: single-stream CTAs for T <= 2 too
            if (a.T <= 2 && !ns1) launch_attention_mma<KV, 2>(a, st);
            else if (attention_stream_wanted(a)) launch_attention_mma_stream<KV>(a, st);
            else launch_attention_mma<KV, 1>(a, st);
            return;
        }
        if (a.T <= 2 && attention_pair_enabled()) { if (a.T == 1) launch_attention_pair<KV, 1>(a, st); else launch_attention_pair<KV, 2>(a, st); return; }
    }
    if (a.T == 1) launch_attention_tq<KV, 1>(a, st);
    else if (a.T == 2) launch_attention_tq<KV, 2>(a, st);
    else if (a.T <= 4) launch_attention_tq<KV, 4>(a, st);
    else launch_attention_tq<KV, 8>(a, st);
}
void launch_attention(const AttnArgs& a, cudaStream_t st) {
    if (a.B <= 0) return;
    if (a.kv_dtype == 0) launch_attention_t<0>(a, st);
    else if (a.kv_dtype == 1) launch_attention_t<1>(a, st);
    else launch_attention_t<2>(a, st);
}

// ------------------------------------------------------------------------------------------
// Conv module core: GLU -> cached causal depthwise conv (k = 9) -> LayerNorm -> SiLU, and the new conv state.
// One CTA per (block of TB <= 8 frames, stream): 256 threads x 4 channels, a 9-deep register window slides over the block's frames.
// The window is primed with the 8 rows before the block: rows of the conv state where they precede the chunk, GLU rows of earlier
// frames recomputed from the pointwise-1 output otherwise (8 x 4 sigmoids per thread, all loads in flight at once) -- a fixed cost
// per block, so TB trades redundancy against the length of the serial chain (launch_conv_module picks it).
// The frame loop holds NO barrier: a frame's conv outputs are parked in shared memory (each thread reads back only its own), its
// LayerNorm statistics are reduced inside the warp and left per warp; ONE block-wide exchange after the loop completes the statistics
// of all frames (two in strict fp32: mean, then centred squares -- the reference's order), then the frames are normalised and stored.
// (The per-frame block reductions were 2-4 barriers per frame on the critical path: 7 frames x 256 streams took 25 us for 34 MB.)
// Per frame the arithmetic is the one of a sequential pass with block_sum_256 / block_mean_var_256: same tap order, same trees.
// The LAST block writes the new state xp[T .. T + 7] (:368-381) into the OTHER parity of the double-buffered state: CTAs of earlier
// blocks may still be reading the old one.
// (Tried and dropped: one CTA per frame -- 9x the GLU work and pointwise-1 reads; 33 us instead of 17 us per layer at 64 x 14 frames.
//  Fully unrolled rows x frames with register accumulators -- 10 k instructions of straight-line code, instruction-fetch bound: +8 us.)
// ------------------------------------------------------------------------------------------
constexpr int CONV_TB = 8;                                                        // frames per CTA, at most
__global__ void __launch_bounds__(256, 2) conv_module_kernel(const ConvModArgs a, int TB) {
    NSB_KERNEL_BEGIN(TR_CONVMOD)                                                  // taps, conv state (written by this layer's kernel of earlier steps only), LN affine: pre-wait
    __shared__ float4 cvs[CONV_TB][256];                                          // conv outputs of the block's frames, [frame][thread]
    __shared__ float red[CONV_TB][16];
    const int t0 = blockIdx.x * TB, b = blockIdx.y, c0 = threadIdx.x * 4, T = a.T;
    const int t1 = min(T, t0 + TB), nt = t1 - t0;
    const int slot = a.slot_of_b[b], par = a.cc_par[slot] & 1;
    const int wrp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool f32 = a.out_type == OUT_F32;
    const float* cache = a.conv_cache + (size_t)slot * a.slot_stride + (size_t)par * a.par_stride;
    float* cache_new = a.conv_cache + (size_t)slot * a.slot_stride + (size_t)(par ^ 1) * a.par_stride;
    float win[4][CONV_K];
    const float4* taps = reinterpret_cast<const float4*>(a.dw_w + c0);          // tap k of the thread's 4 channels: taps[k * 256]; re-read per frame (L1), not held in registers
#pragma unroll
    for (int k = 0; k < CONV_K - 1; ++k) {                                       // window rows xp[t0 + k] that lie in the state: t0 + k < 8   (xp = [state(8) || glu(T)] :323-328)
        if (t0 + k < CONV_K - 1) {
            const float4 v = *(const float4*)(cache + (size_t)(t0 + k) * D_MODEL + c0);
            win[0][k] = v.x; win[1][k] = v.y; win[2][k] = v.z; win[3][k] = v.w;
        }
    }
    NSB_KERNEL_WAIT()
    const float* row0 = a.pw1 + (size_t)b * T * 2 * D_MODEL;
    if (t0 > 0) {                                                                // the other window rows: GLU of frames t0 + k - 8 >= 0, recomputed (block-uniform branch)
#pragma unroll
        for (int k0 = 0; k0 < CONV_K - 1; k0 += 4) {                             // four rows' loads in flight at a time (registers)
            float4 ha[4], hg[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int r = t0 + k0 + k - (CONV_K - 1);
                if (r >= 0) {
                    const float* rp = row0 + (size_t)r * 2 * D_MODEL;
                    ha[k] = ld4_planes<PW1_MAX_PLANES>(rp + c0, a.planes, a.plane_stride); hg[k] = ld4_planes<PW1_MAX_PLANES>(rp + D_MODEL + c0, a.planes, a.plane_stride);
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (t0 + k0 + k - (CONV_K - 1) >= 0) {
                    win[0][k0 + k] = ha[k].x * sigmoid_exact(hg[k].x); win[1][k0 + k] = ha[k].y * sigmoid_exact(hg[k].y);
                    win[2][k0 + k] = ha[k].z * sigmoid_exact(hg[k].z); win[3][k0 + k] = ha[k].w * sigmoid_exact(hg[k].w);
                }
            }
        }
    }
    // rows t, t + 1, t + 2 of the block in flight: the loop holds no barrier, so the loads of a warp run ahead of its arithmetic
    auto row_a = [&](int t) { return ld4_planes<PW1_MAX_PLANES>(row0 + (size_t)(t0 + t) * 2 * D_MODEL + c0, a.planes, a.plane_stride); };
    auto row_g = [&](int t) { return ld4_planes<PW1_MAX_PLANES>(row0 + (size_t)(t0 + t) * 2 * D_MODEL + D_MODEL + c0, a.planes, a.plane_stride); };
    float4 av = row_a(0), gv = row_g(0), av1 = av, gv1 = gv, av2 = av, gv2 = gv;
    if (1 < nt) { av1 = row_a(1); gv1 = row_g(1); }
    if (2 < nt) { av2 = row_a(2); gv2 = row_g(2); }
    for (int t = 0; t < nt; ++t) {
        float4 av3 = av2, gv3 = gv2;
        if (t + 3 < nt) { av3 = row_a(t + 3); gv3 = row_g(t + 3); }
        win[0][8] = av.x * sigmoid_exact(gv.x); win[1][8] = av.y * sigmoid_exact(gv.y);   // GLU :629-636
        win[2][8] = av.z * sigmoid_exact(gv.z); win[3][8] = av.w * sigmoid_exact(gv.w);
        float cv[4];
#pragma unroll
        for (int k = 0; k < CONV_K; ++k) {                                       // :341-360, taps in ascending order
            const float4 w4 = __ldg(taps + k * (D_MODEL / 4));
            if (k == 0) { cv[0] = win[0][0] * w4.x; cv[1] = win[1][0] * w4.y; cv[2] = win[2][0] * w4.z; cv[3] = win[3][0] * w4.w; }
            else { cv[0] = fmaf(win[0][k], w4.x, cv[0]); cv[1] = fmaf(win[1][k], w4.y, cv[1]); cv[2] = fmaf(win[2][k], w4.z, cv[2]); cv[3] = fmaf(win[3][k], w4.w, cv[3]); }
        }
        cvs[t][threadIdx.x] = make_float4(cv[0], cv[1], cv[2], cv[3]);
        if (f32) {                                                               // strict fp32, pass 1: the warp's part of the row sum (block_sum_256's tree)
            const float s = warp_sum(cv[0] + cv[1] + cv[2] + cv[3]);
            if (lane == 0) red[t][wrp] = s;
        } else {                                                                 // 16-bit modes: the warp's (mean, M2) by Chan's pairwise merge (block_mean_var_256's tree)
            float m = ((cv[0] + cv[1]) + (cv[2] + cv[3])) * 0.25f;
            float M2 = (cv[0] - m) * (cv[0] - m) + (cv[1] - m) * (cv[1] - m) + (cv[2] - m) * (cv[2] - m) + (cv[3] - m) * (cv[3] - m);
            float half_n = 2.0f;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const float mo = __shfl_xor_sync(0xffffffffu, m, o), M2o = __shfl_xor_sync(0xffffffffu, M2, o);
                const float d = mo - m;
                M2 = (M2 + M2o) + d * d * half_n; m = m + 0.5f * d; half_n *= 2.0f;
            }
            if (lane == 0) { red[t][2 * wrp] = m; red[t][2 * wrp + 1] = M2; }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int k = 0; k < CONV_K - 1; ++k) win[u][k] = win[u][k + 1];
        av = av1; gv = gv1; av1 = av2; gv1 = gv2; av2 = av3; gv2 = gv3;
    }
    if (t1 == T) {                                                               // new state = last 8 rows of xp :368-381 (the window after the last frame)
#pragma unroll
        for (int k = 0; k < CONV_K - 1; ++k)
            *(float4*)(cache_new + (size_t)k * D_MODEL + c0) = make_float4(win[0][k], win[1][k], win[2][k], win[3][k]);
    }
    const float4 g4 = *(const float4*)(a.ln_g + c0), b4 = *(const float4*)(a.ln_b + c0);
    __syncthreads();                                                             // every warp's partial statistics of every frame
    if (f32) {                                                                   // pass 2: centred squares against the complete mean
        for (int t = 0; t < nt; ++t) {
            float tot = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) tot += red[t][i];
            const float mean = tot * (1.0f / D_MODEL);
            const float4 c = cvs[t][threadIdx.x];
            const float e0 = c.x - mean, e1 = c.y - mean, e2 = c.z - mean, e3 = c.w - mean;
            const float s = warp_sum(e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3);
            if (lane == 0) red[t][8 + wrp] = s;
        }
        __syncthreads();
    }
    for (int t = 0; t < nt; ++t) {                                               // LN :643-645, SiLU :646
        float mean, var;
        if (f32) {
            float tot = 0.f, tq = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) { tot += red[t][i]; tq += red[t][8 + i]; }
            mean = tot * (1.0f / D_MODEL); var = tq * (1.0f / D_MODEL);
        } else {
            float mm[8], MM[8], hn = 64.0f;                                      // n / 2 of the groups being merged: 8 warps of 128 values -> 4 x 256 -> 2 x 512 -> 1024
#pragma unroll
            for (int i = 0; i < 8; ++i) { mm[i] = red[t][2 * i]; MM[i] = red[t][2 * i + 1]; }
#pragma unroll
            for (int n = 8; n > 1; n >>= 1) {
#pragma unroll
                for (int i = 0; i < n / 2; ++i) {
                    const float d = mm[2 * i + 1] - mm[2 * i];
                    MM[i] = (MM[2 * i] + MM[2 * i + 1]) + d * d * hn; mm[i] = mm[2 * i] + 0.5f * d;
                }
                hn *= 2.0f;
            }
            mean = mm[0]; var = MM[0] * (1.0f / D_MODEL);
        }
        const float4 c = cvs[t][threadIdx.x];
        const float rs = 1.0f / sqrtf(var + 1e-5f);
        const float4 y = make_float4(silu_exact((c.x - mean) * rs * g4.x + b4.x), silu_exact((c.y - mean) * rs * g4.y + b4.y),
                                     silu_exact((c.z - mean) * rs * g4.z + b4.z), silu_exact((c.w - mean) * rs * g4.w + b4.w));
        store4_out(a.out, ((size_t)b * T + t0 + t) * D_MODEL + c0, y, a.out_type);
    }
    NSB_KERNEL_EPILOGUE();
}
void launch_conv_module(const ConvModArgs& a, cudaStream_t st) {
    static const int tb_env = [] { const char* e = getenv("NSB_CONV_TB"); return e ? atoi(e) : 0; }();
    // Frames per block: a block primes its window with 8 recomputed rows, so few blocks are cheapest in work, but the kernel is bound by
    // its load -> GLU -> conv -> reduce chain, not by work: aim for ~256 CTAs, never less than 2 frames per block. Measured per step:
    // 32 streams x 7 frames 2.57 (one block) -> 2.46 ms (blocks of 2), 64 x 7 3.29 -> 3.21 (blocks of 3), 64 x 14 5.09 (blocks of 7) ->
    // 5.01 (blocks of 4); 256 x 7 stays one block per stream (6.98 vs 7.14 ms with two), the batch path (one stream x ~2000 frames) 8-frame blocks.
    int blocks = std::max(1, std::min((256 + a.B - 1) / std::max(a.B, 1), (a.T + 1) / 2));
    int TB = (a.T + blocks - 1) / blocks;
    if (tb_env > 0) TB = tb_env;
    TB = std::max(1, std::min(TB, CONV_TB));
    if (a.B > 0 && a.T > 0) launch_k(conv_module_kernel, dim3((a.T + TB - 1) / TB, a.B), dim3(256), 0, st, a, TB);
}

__global__ void advance_streams_kernel(const int* __restrict__ slot_of_b, int B, int T, int* ring_pos, int* valid_len, int* cc_par) {
    NSB_KERNEL_PROLOGUE(TR_ADVANCE)
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int s = slot_of_b[b];
    ring_pos[s] = (ring_pos[s] + T) % (ATT_L + T);
    valid_len[s] = min(valid_len[s] + T, ATT_L);                                 // :1018
    cc_par[s] ^= 1;                                                              // every layer's conv module wrote the other parity of the conv state
    NSB_KERNEL_EPILOGUE();
}
void launch_advance_streams(const int* slot_of_b, int B, int T, int* ring_pos, int* valid_len, int* cc_par, cudaStream_t st) {
    if (B > 0) launch_k(advance_streams_kernel, dim3((B + 127) / 128), dim3(128), 0, st, slot_of_b, B, T, ring_pos, valid_len, cc_par);
}

// ------------------------------------------------------------------------------------------
// Full-context rel-pos attention of the non-streaming batch path (build_rel_pos_mha + build_rel_shift, nemo-ggml.cpp:548-680).
// Correctness-first SIMT (validated on hardware, tests/test_zz_batch_path.py; the path is bound by the single-stream greedy decode, not by
// this kernel: DESIGN.md section 1). One warp = one (head, query row i), lane l owns head dims 4l .. 4l+3.
//   pass 1: S[j] = ((q_i + u).k_j + (q_i + v).P[i - j]) / sqrt(128) for all T keys into shared memory -- the pad / reshape / drop
//           rel-shift is index arithmetic: BD[i, j] = BD_raw[i, j + T - 1 - i] <-> relative position i - j;
//   pass 2: softmax over the row (max-subtract, expf, sum), as ggml_soft_max;
//   pass 3: ctx_i = sum_j p_j v_j in key order.
// K, V and P are read at the K/V dtype's precision, as the streaming ring stores them.
// ------------------------------------------------------------------------------------------
constexpr int ATTF_WARPS = 4;
__global__ void __launch_bounds__(32 * ATTF_WARPS) attention_full_kernel(const AttnFullArgs a) {
    extern __shared__ float attf_sc[];                                           // [ATTF_WARPS][T]
    NSB_KERNEL_PROLOGUE(TR_ATTN)
    const int h = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, T = a.T, kv = a.kv_dtype;
    const int i = blockIdx.y * ATTF_WARPS + warp;
    if (i < T) {                                                                 // warp-uniform; no block-wide barrier below
        float* sc = attf_sc + (size_t)warp * T;
        const int c = h * D_HEAD + lane * 4;
        const float4 q4 = *(const float4*)(a.qkv + (size_t)i * 3 * D_MODEL + c);
        const float4 u4 = *(const float4*)(a.bias_u + c), w4 = *(const float4*)(a.bias_v + c);
        const float qu[4] = {q4.x + u4.x, q4.y + u4.y, q4.z + u4.z, q4.w + u4.w};
        const float qv[4] = {q4.x + w4.x, q4.y + w4.y, q4.z + w4.z, q4.w + w4.w};
        const float scale = 1.0f / sqrtf((float)D_HEAD);
        for (int j = 0; j < T; ++j) {
            const float4 k4 = *(const float4*)(a.qkv + (size_t)j * 3 * D_MODEL + D_MODEL + c);
            const size_t pr = (size_t)(i - j + a.pos_center) * D_MODEL + c;
            float ac = round_kv(k4.x, kv) * qu[0];
            ac = fmaf(round_kv(k4.y, kv), qu[1], ac); ac = fmaf(round_kv(k4.z, kv), qu[2], ac); ac = fmaf(round_kv(k4.w, kv), qu[3], ac);
            float bd = load_kv(a.pos_proj, pr, kv) * qv[0];
            bd = fmaf(load_kv(a.pos_proj, pr + 1, kv), qv[1], bd); bd = fmaf(load_kv(a.pos_proj, pr + 2, kv), qv[2], bd);
            bd = fmaf(load_kv(a.pos_proj, pr + 3, kv), qv[3], bd);
            ac = warp_sum(ac); bd = warp_sum(bd);
            if (lane == 0) sc[j] = (ac + bd) * scale;
        }
        __syncwarp();
        float mx = -INFINITY;
        for (int j = lane; j < T; j += 32) mx = fmaxf(mx, sc[j]);
        mx = warp_max(mx);
        float sum = 0.0f;
        for (int j = lane; j < T; j += 32) { const float e = expf(sc[j] - mx); sc[j] = e; sum += e; }
        sum = warp_sum(sum);
        __syncwarp();
        const float inv = 1.0f / sum;
        float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        for (int j = 0; j < T; ++j) {
            const float p = sc[j] * inv;
            const float4 v4 = *(const float4*)(a.qkv + (size_t)j * 3 * D_MODEL + 2 * D_MODEL + c);
            acc[0] = fmaf(round_kv(v4.x, kv), p, acc[0]); acc[1] = fmaf(round_kv(v4.y, kv), p, acc[1]);
            acc[2] = fmaf(round_kv(v4.z, kv), p, acc[2]); acc[3] = fmaf(round_kv(v4.w, kv), p, acc[3]);
        }
        const size_t o = (size_t)i * D_MODEL + c;
#pragma unroll
        for (int u = 0; u < 4; ++u) store_out(a.ctx, o + u, acc[u], a.out_type);
    }
    NSB_KERNEL_EPILOGUE();
}
void launch_attention_full(const AttnFullArgs& a, cudaStream_t st) {
    if (a.T < 1 || a.T > 2048) throw CudaError("attention_full: 1 <= frames <= 2048 (the reference's positional table, nemo-ggml.cpp:196)");
    launch_k(attention_full_kernel, dim3(N_HEADS, (a.T + ATTF_WARPS - 1) / ATTF_WARPS), dim3(32 * ATTF_WARPS), (size_t)ATTF_WARPS * a.T * sizeof(float), st, a);
}

}  // namespace nsb
