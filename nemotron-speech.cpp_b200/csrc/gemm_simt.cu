// gemm_simt.cu -- fp32 SIMT GEMM  C[M,N] = A[M,K] * W[N,K]^T  with fused epilogues.
//
// Used where the reference computes in f32 (ggml_mul_mat with F32 src0): the subsampling stem's 1x1 convs and
// out-projection (src/nemo-ggml.cpp:906-946), joint.enc (:1080-1081), and -- in NSB_COMPUTE_F32 strict-parity
// mode -- the per-layer matrices too. The 16-bit / Q8_0 paths use the tcgen05 kernel in gemm_tc.cu.
//
// Both operands are K-contiguous. 64x64x16 tiles, 256 threads, 4x4 register micro-tiles, operands staged
// transposed in shared memory; global loads are float4 along K (coalesced 64 B per row segment).
#include "kernels.cuh"

namespace nsb {

NSB_DEFINE_TRACE_BINDER(trace_bind_simt)

namespace {
constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;

__device__ __forceinline__ const float* a_row_ptr(const GemmArgs& g, int m) {
    const float* A = (const float*)g.A;
    if (g.group == 0) return A + (size_t)m * g.lda;
    return A + (size_t)(m / g.group) * g.group_stride + (size_t)(m % g.group + g.row_off) * g.lda;
}

__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmArgs g) {
    NSB_KERNEL_PROLOGUE(TR_GEMM_SIMT)
    __shared__ float As[2][BK][BM + PAD], Bs[2][BK][BN + PAD];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int lr = tid >> 2, lk = (tid & 3) * 4;                       // loader: row 0..63, k offset 0,4,8,12
    const bool a_ok = (m0 + lr) < g.M, b_ok = (n0 + lr) < g.N;
    const float* a_ptr = a_ok ? a_row_ptr(g, m0 + lr) + lk : nullptr;
    const float* b_ptr = b_ok ? (const float*)g.W + (size_t)(n0 + lr) * g.K + lk : nullptr;
    float acc[4][4] = {};
    const int nk = g.K / BK;
    float4 ra = a_ok ? *(const float4*)a_ptr : make_float4(0, 0, 0, 0);
    float4 rb = b_ok ? *(const float4*)b_ptr : make_float4(0, 0, 0, 0);
    for (int kt = 0; kt < nk; ++kt) {
        const int s = kt & 1;
        As[s][lk + 0][lr] = ra.x; As[s][lk + 1][lr] = ra.y; As[s][lk + 2][lr] = ra.z; As[s][lk + 3][lr] = ra.w;
        Bs[s][lk + 0][lr] = rb.x; Bs[s][lk + 1][lr] = rb.y; Bs[s][lk + 2][lr] = rb.z; Bs[s][lk + 3][lr] = rb.w;
        __syncthreads();
        if (kt + 1 < nk) {                                             // prefetch next k-tile into registers
            ra = a_ok ? *(const float4*)(a_ptr + (size_t)(kt + 1) * BK) : make_float4(0, 0, 0, 0);
            rb = b_ok ? *(const float4*)(b_ptr + (size_t)(kt + 1) * BK) : make_float4(0, 0, 0, 0);
        }
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a4 = *(const float4*)&As[s][k][ty * 4];
            const float4 b4 = *(const float4*)&Bs[s][k][tx * 4];
            const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        // double-buffered smem: the next iteration writes the other stage, so one barrier per k-tile suffices
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= g.N) continue;
            float v = acc[i][j];
            if (g.bias) v += g.bias[n];
            const size_t o = (size_t)m * g.ldc + n;
            if (g.epi == EPI_RELU) v = fmaxf(v, 0.0f);
            else if (g.epi == EPI_SILU) v = silu_exact(v);
            else if (g.epi == EPI_RESID) { float* C = (float*)g.C; C[o] = C[o] + g.alpha * v; continue; }
            store_out(g.C, o, v, g.out_type);
        }
    }
    NSB_KERNEL_EPILOGUE();
}
}  // namespace

void launch_gemm_simt(const GemmArgs& a, cudaStream_t st) {
    if (a.M <= 0 || a.N <= 0) return;
    if (a.K % BK != 0 || (a.lda % 4) != 0 || (a.group_stride % 4) != 0)
        throw CudaError("gemm_simt: K must be a multiple of 16 and rows 16-byte aligned");
    dim3 grid((a.N + BN - 1) / BN, (a.M + BM - 1) / BM);
    launch_k(gemm_simt_kernel, grid, dim3(256), 0, st, a);
}

}  // namespace nsb
