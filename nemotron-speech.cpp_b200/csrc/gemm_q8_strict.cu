// gemm_q8_strict.cu -- Q8_0 x Q8_0 GEMM with the REFERENCE's arithmetic (NSB_COMPUTE_Q8_0_STRICT), for sm_100a.
//
// What the reference does for a Q8_0 weight: ggml_mul_mat (src/nemo-stream.cpp:457-459, :488, :542, :571-573, :626, :649) quantises
// every ACTIVATION row to Q8_0 as well (upstream ggml quantize_row_q8_0: per 32 values d = amax / 127, q = roundf(x * (1 / d)),
// d stored as fp16) and takes integer block dot products (vec_dot_q8_0_q8_0):
//     y[r][o] = sum_b  float(sum_i qw[o][32 b + i] * qx[r][32 b + i]) * (fp16(d_w[o][b]) * fp16(d_x[r][b]))      in f32, b ascending.
// The fast Q8_0 mode (gemm_tc.cu: gemm_q8_kernel) keeps activations in fp16 -- more accurate than the reference, hence not
// comparable bit for bit. This file is the strict mode:
//   * quantize_rows_q8_kernel: the activation quantiser, operation for operation (IEEE division, reciprocal, roundf);
//   * gemm_q8_strict_kernel:   one mma.sync.m16n8k32 (s8 x s8 -> s32, exact) per 32-value block = exactly one Q8_0 block dot per
//                              output element, then the scale product and the accumulation in f32 with separate multiply and add
//                              (the reference build contracts nothing), blocks in ascending order.
// Every output element therefore goes through the same sequence of roundings as the scalar reference loop: results are
// bit-identical to the CPU checker's restatement of that loop (the test suite asserts equality of the bit patterns), not merely close.
// A parity mode like NSB_COMPUTE_F32 (no split-K, no tcgen05: the per-block rescale needs the integer sums in registers).
#include "kernels.cuh"

namespace nsb {

namespace {

// 8 lanes per 32-value block (one float4 each); blocks are consecutive along the row, rows consecutive in `q` / `d`.
__global__ void __launch_bounds__(256) quantize_rows_q8_kernel(const float* __restrict__ A, long long lda, int M, int K,
                                                               int8_t* __restrict__ q, float* __restrict__ d) {
    NSB_KERNEL_PROLOGUE(TR_OTHER)
    const int nb = K / 32;
    const long long n_blocks = (long long)M * nb;
    for (long long blk = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3; blk < ((n_blocks + 31) & ~31LL); blk += ((long long)gridDim.x * blockDim.x) >> 3) {
        const bool live = blk < n_blocks;                           // whole warps stay in the loop: the shuffles below are warp-wide
        const long long r = live ? blk / nb : 0; const int b = live ? (int)(blk % nb) : 0, l8 = threadIdx.x & 7;
        const float4 v = live ? *reinterpret_cast<const float4*>(A + r * lda + b * 32 + l8 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        float amax = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        const float dd = __fdiv_rn(amax, 127.0f);
        const float id = dd != 0.0f ? __fdiv_rn(1.0f, dd) : 0.0f;
        if (!live) continue;
        char4 c;
        c.x = (signed char)roundf(__fmul_rn(v.x, id)); c.y = (signed char)roundf(__fmul_rn(v.y, id));
        c.z = (signed char)roundf(__fmul_rn(v.z, id)); c.w = (signed char)roundf(__fmul_rn(v.w, id));
        *reinterpret_cast<char4*>(q + r * K + b * 32 + l8 * 4) = c;
        if (l8 == 0) d[r * nb + b] = __half2float(__float2half_rn(dd));
    }
    NSB_KERNEL_EPILOGUE();
}

__device__ __forceinline__ void imma_16832(int (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
                 : "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "r"(0), "r"(0), "r"(0), "r"(0));
}

struct Q8sArgs {
    const int8_t* qa; const float* da;        // quantised activations [M][K], scales [M][K/32] (fp16 values widened)
    const int8_t* qw; const __half* dw;       // weight planes [N][K], [N][K/32]
    int M, N, K;
    float* C; long long ldc; int epi; float alpha;
};

constexpr int SBM = 64, SBN = 64;             // CTA tile; 4 warps as 2 x 2, a warp owns 32 x 32 = 2 (m16) x 4 (n8) MMA tiles

struct Frag { uint32_t a[2][4]; uint32_t b[4][2]; float da[2][2]; float dw[4][2]; };

__device__ __forceinline__ void load_frag(Frag& f, const Q8sArgs& p, const int (&arow)[2][2], const int (&bcol)[4], const int (&ccol)[4][2], int blk, int tq) {
    const int nb = p.K / 32;
    const size_t ko = (size_t)blk * 32 + 4 * tq;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
        const int8_t* r0 = p.qa + (size_t)arow[mi][0] * p.K + ko; const int8_t* r1 = p.qa + (size_t)arow[mi][1] * p.K + ko;
        f.a[mi][0] = *reinterpret_cast<const uint32_t*>(r0); f.a[mi][1] = *reinterpret_cast<const uint32_t*>(r1);
        f.a[mi][2] = *reinterpret_cast<const uint32_t*>(r0 + 16); f.a[mi][3] = *reinterpret_cast<const uint32_t*>(r1 + 16);
        f.da[mi][0] = p.da[(size_t)arow[mi][0] * nb + blk]; f.da[mi][1] = p.da[(size_t)arow[mi][1] * nb + blk];
    }
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
        const int8_t* w = p.qw + (size_t)bcol[ni] * p.K + ko;
        f.b[ni][0] = *reinterpret_cast<const uint32_t*>(w); f.b[ni][1] = *reinterpret_cast<const uint32_t*>(w + 16);
        f.dw[ni][0] = __half2float(p.dw[(size_t)ccol[ni][0] * nb + blk]); f.dw[ni][1] = __half2float(p.dw[(size_t)ccol[ni][1] * nb + blk]);
    }
}

__global__ void __launch_bounds__(128) gemm_q8_strict_kernel(const Q8sArgs p) {
    NSB_KERNEL_PROLOGUE(TR_OTHER)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
    const int m0 = blockIdx.y * SBM + (warp >> 1) * 32, n0 = blockIdx.x * SBN + (warp & 1) * 32;
    // fragment coordinates (mma.m16n8k32: A row g / g + 8, B column g, C rows g / g + 8 x columns 2 tq / 2 tq + 1); rows past M are
    // clamped for the loads and never stored
    int arow[2][2], bcol[4], ccol[4][2];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) { arow[mi][0] = min(m0 + mi * 16 + g, p.M - 1); arow[mi][1] = min(m0 + mi * 16 + g + 8, p.M - 1); }
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) { bcol[ni] = n0 + ni * 8 + g; ccol[ni][0] = n0 + ni * 8 + 2 * tq; ccol[ni][1] = ccol[ni][0] + 1; }
    // A-side scales are needed per C row (g, g + 8): the same rows the A fragment uses. W-side scales per C column (2 tq, 2 tq + 1).
    float acc[2][4][4];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[mi][ni][e] = 0.0f;
    const int nb = p.K / 32;
    Frag cur, nxt;
    load_frag(cur, p, arow, bcol, ccol, 0, tq);
    for (int blk = 0; blk < nb; ++blk) {
        if (blk + 1 < nb) load_frag(nxt, p, arow, bcol, ccol, blk + 1, tq);     // next block's operands in flight during this block's MMAs
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                int c[4];
                imma_16832(c, cur.a[mi], cur.b[ni]);
                // sumf += (float)sumi * (d_w * d_x): multiply and add rounded separately, as in the reference's scalar loop
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float sc = __fmul_rn(cur.dw[ni][e & 1], cur.da[mi][e >> 1]);
                    acc[mi][ni][e] = __fadd_rn(acc[mi][ni][e], __fmul_rn((float)c[e], sc));
                }
            }
        cur = nxt;
    }
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int e2 = 0; e2 < 2; ++e2) {
            const int row = m0 + mi * 16 + g + 8 * e2;
            if (row >= p.M) continue;
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                float2 v = make_float2(acc[mi][ni][2 * e2], acc[mi][ni][2 * e2 + 1]);
                float* dst = p.C + (size_t)row * p.ldc + ccol[ni][0];
                if (p.epi == EPI_SILU) { v.x = silu_exact(v.x); v.y = silu_exact(v.y); }
                else if (p.epi == EPI_RESID) {
                    const float2 x = *reinterpret_cast<const float2*>(dst);
                    v.x = __fadd_rn(x.x, __fmul_rn(p.alpha, v.x)); v.y = __fadd_rn(x.y, __fmul_rn(p.alpha, v.y));
                }
                *reinterpret_cast<float2*>(dst) = v;
            }
        }
    NSB_KERNEL_EPILOGUE();
}
}  // namespace

size_t q8_strict_scratch_bytes(int rows, int K) { return (size_t)rows * K + (size_t)rows * (K / 32) * 4 + 256; }

// a.A: f32 activations [M][K] (lda elements); a.W / a.w_scales: Q8_0 planes; a.C f32; epilogues NONE / SILU / RESID, no bias
void launch_gemm_q8_strict(const GemmArgs& a, void* scratch, size_t scratch_bytes, cudaStream_t st) {
    if (a.M <= 0) return;
    if (!a.w_scales) throw CudaError("gemm_q8_strict: the weight has no Q8_0 planes");
    if (a.K % 32 != 0 || a.N % SBN != 0 || (a.lda % 4) != 0 || (a.ldc % 2) != 0 || a.group != 0 || a.bias || a.out_type != OUT_F32 ||
        !(a.epi == EPI_NONE || a.epi == EPI_SILU || a.epi == EPI_RESID))
        throw CudaError("gemm_q8_strict: unsupported shape / epilogue");
    if (q8_strict_scratch_bytes(a.M, a.K) > scratch_bytes) throw CudaError("gemm_q8_strict: activation scratch too small");
    int8_t* qa = (int8_t*)scratch;
    float* da = (float*)((char*)scratch + (((size_t)a.M * a.K + 255) & ~(size_t)255));
    const long long n_thr = (long long)a.M * (a.K / 32) * 8;
    const int blocks = (int)std::min<long long>((n_thr + 255) / 256, 148 * 16);
    launch_k(quantize_rows_q8_kernel, dim3(blocks), dim3(256), 0, st, (const float*)a.A, a.lda, a.M, a.K, qa, da);
    Q8sArgs p{qa, da, (const int8_t*)a.W, (const __half*)a.w_scales, a.M, a.N, a.K, (float*)a.C, a.ldc, a.epi, a.alpha};
    launch_k(gemm_q8_strict_kernel, dim3(a.N / SBN, (a.M + SBM - 1) / SBM), dim3(128), 0, st, p);
}

}  // namespace nsb
