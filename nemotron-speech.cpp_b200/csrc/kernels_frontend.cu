// kernels_frontend.cu -- log-mel front-end and dw-striding subsampling stem (sm_100a).
//
// Reference behaviour being replaced:
//   * nemo_preprocessor_process / stft_magnitude / fft_frame  (src/preprocessor.cpp:330-395,167-205,113-161)
//   * build_causal_conv2d / build_causal_dw_conv2d / build_conv_subsampling stem (src/nemo-ggml.cpp:820-930)
// The reference runs the mel on one CPU thread per context; here one CTA computes one frame of one
// stream, so a step launches (8T x B) CTAs.
#include "kernels.cuh"

namespace nsb {

NSB_DEFINE_TRACE_BINDER(trace_bind_frontend)

// ------------------------------------------------------------------------------------------
// log-mel: one CTA = one 512-sample frame. Shared-memory radix-2 DIT FFT (9 stages x 256
// butterflies), power spectrum, 128x257 filterbank, log.
// Arithmetic is written with explicit _rn intrinsics (no FMA contraction) in the same operation
// order as the reference so that every step up to the final log is bit-identical to its x86 build;
// the log is taken in fp64 and rounded once (glibc logf is correctly rounded in nearly all cases).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) logmel_kernel(const int16_t* __restrict__ pcm, int pcm_row_stride, int n_frames,
                                                     const float* __restrict__ window, const float* __restrict__ cos_t,
                                                     const float* __restrict__ sin_t, const float* __restrict__ fb_t,
                                                     float* __restrict__ mel_out, size_t out_batch_stride) {
    NSB_KERNEL_PROLOGUE(TR_LOGMEL)
    __shared__ float re[N_FFT], im[N_FFT], pw[N_BINS + 3];
    const int j = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
    const int16_t* row = pcm + (size_t)b * pcm_row_stride + (size_t)j * HOP;   // row[0] = sample before the frame
    const float scale = 1.0f / 32768.0f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int i = t + h * 256;
        const float cur = (float)row[1 + i] * scale, prev = (float)row[i] * scale;     // exact
        const float y = __fsub_rn(cur, __fmul_rn(0.97f, prev));                          // pre-emphasis :349-356
        const int r = __brev((unsigned)i) >> 23;                                         // 9-bit reversal :124-127
        re[r] = __fmul_rn(y, window[i]);                                                 // :184-194
        im[r] = 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int m2 = 1; m2 < N_FFT; m2 <<= 1) {                                             // :131-154
        const int jj = t & (m2 - 1), i1 = ((t - jj) << 1) + jj, i2 = i1 + m2;
        const int idx = jj * (N_FFT / (2 * m2));
        const float wr = cos_t[idx], wi = -sin_t[idx];
        const float r2 = re[i2], q2 = im[i2], r1 = re[i1], q1 = im[i1];
        const float tr = __fsub_rn(__fmul_rn(wr, r2), __fmul_rn(wi, q2));
        const float ti = __fadd_rn(__fmul_rn(wr, q2), __fmul_rn(wi, r2));
        re[i2] = __fsub_rn(r1, tr); im[i2] = __fsub_rn(q1, ti);
        re[i1] = __fadd_rn(r1, tr); im[i1] = __fadd_rn(q1, ti);
        __syncthreads();
    }
    for (int k = t; k < N_BINS; k += 256) {                                              // :201, :363-368
        const float mag = __fsqrt_rn(__fadd_rn(__fmul_rn(re[k], re[k]), __fmul_rn(im[k], im[k])));
        pw[k] = __fmul_rn(mag, mag);
    }
    __syncthreads();
    if (t < N_MELS) {                                                                    // :374-383
        // the filterbank rows are (nearly) banded: [lo, hi) = span of the non-zero weights of mel bin t, found at load time.
        // Terms outside add +/-0 to a non-negative partial sum, so skipping them is bit-exact -- and takes the 131 KB table
        // read per CTA down to the few KB that matter.
        const int* rng = reinterpret_cast<const int*>(fb_t + N_BINS * N_MELS);
        const int lo = rng[t], hi = rng[N_MELS + t];
        float sum = 0.0f;
#pragma unroll 4
        for (int k = lo; k < hi; ++k) sum = __fadd_rn(sum, __fmul_rn(fb_t[k * N_MELS + t], pw[k]));
        const float v = __fadd_rn(sum, 5.960464477539063e-8f);
        mel_out[(size_t)b * out_batch_stride + (size_t)j * N_MELS + t] = (float)log((double)v);
    }
    NSB_KERNEL_EPILOGUE();
}

void launch_logmel(const int16_t* pcm, int pcm_row_stride, int B, int n_frames, const float* window512, const float* cos_t,
                   const float* sin_t, const float* fb_t, float* mel_out, size_t out_batch_stride, cudaStream_t st) {
    if (B <= 0 || n_frames <= 0) return;
    launch_k(logmel_kernel, dim3(n_frames, B), dim3(256), 0, st, pcm, pcm_row_stride, n_frames, window512, cos_t, sin_t, fb_t, mel_out,
             out_batch_stride);
}

// ------------------------------------------------------------------------------------------
// chunk image access: frame f of batch row b  (f < 9 -> history of the slot, else new frames)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float chunk_mel(const float* __restrict__ hist, const float* __restrict__ mel_new, int slot, int b,
                                           int T, int f, int m) {
    return f < PRE_CACHE ? hist[((size_t)slot * PRE_CACHE + f) * N_MELS + m]
                         : mel_new[((size_t)b * 8 * T + (f - PRE_CACHE)) * N_MELS + m];
}

// conv0: Cin = 1 -> 256, 3x3 stride 2, pad (2,1) on time and freq, + bias, ReLU. One CTA = one output pixel,
// thread = output channel (NHWC store is fully coalesced; the 9 taps are warp-broadcast loads).
__global__ void __launch_bounds__(SUB_CH) conv0_kernel(const float* __restrict__ hist, const float* __restrict__ mel_new,
                                                       const int* __restrict__ slot_of_b, int T, const float* __restrict__ w_t,
                                                       const float* __restrict__ bias, float* __restrict__ out) {
    NSB_KERNEL_PROLOGUE(TR_OTHER)
    const int ow = blockIdx.x, oh = blockIdx.y, b = blockIdx.z, oc = threadIdx.x;
    const int M = PRE_CACHE + 8 * T, t1 = gridDim.y, W1 = gridDim.x;
    const int slot = slot_of_b[b];
    float acc = 0.0f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
        const int ih = 2 * oh + kh - 2;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
            const int iw = 2 * ow + kw - 2;
            const float x = (ih >= 0 && ih < M && iw >= 0 && iw < N_MELS) ? chunk_mel(hist, mel_new, slot, b, T, ih, iw) : 0.0f;
            acc = fmaf(x, w_t[(kh * 3 + kw) * SUB_CH + oc], acc);
        }
    }
    acc += bias[oc];
    out[(((size_t)b * t1 + oh) * W1 + ow) * SUB_CH + oc] = fmaxf(acc, 0.0f);
    NSB_KERNEL_EPILOGUE();
}

void launch_conv0(const float* mel_hist, const float* mel_new, const int* slot_of_b, int B, int T, const float* w_t,
                  const float* bias, float* out, cudaStream_t st) {
    const int M = PRE_CACHE + 8 * T, t1 = M / 2 + 1, W1 = N_MELS / 2 + 1;
    launch_k(conv0_kernel, dim3(W1, t1, B), dim3(SUB_CH), 0, st, mel_hist, mel_new, slot_of_b, T, w_t, bias, out);
}

// conv0 (+bias, ReLU) fused into the first depthwise 3x3 s2 conv (+bias): the [t1][65][256] conv0 image (55 MB per
// 64-stream step) never goes to HBM. One CTA = one output row (b, oh2); thread = channel; the 7 mel rows the row depends
// on sit in shared memory (broadcast reads); a 3x3 register window of conv0 values slides along the frequency axis, so
// every conv0 value is computed once per output row. Same accumulation order as the two stand-alone kernels (bit-identical).
// FULL = the non-streaming batch path (nemo_encode, nemo-ggml.cpp:961-1003): `hist` is a flat mel image [B][M][128] and `T` carries
// M itself (any length; the chunk variant derives M = 9 + 8T and reads [9 history frames | 8T new frames]).
template <bool FULL>
__global__ void __launch_bounds__(SUB_CH) stem_conv0_dw_kernel(const float* __restrict__ hist, const float* __restrict__ mel_new,
                                                               const int* __restrict__ slot_of_b, int T, const float* __restrict__ w0_t,
                                                               const float* __restrict__ b0, const float* __restrict__ w2_t,
                                                               const float* __restrict__ b2, float* __restrict__ out, int split) {
    // mel column iw lives at index iw + 4: the five columns 4 ow2 - 4 .. 4 ow2 one output column needs start 16-byte aligned
    // (one LDS.128 + one LDS.32 per row instead of 18 scalar broadcast loads: the kernel was bound by shared-memory issue)
    constexpr int MS = N_MELS + 8;
    __shared__ __align__(16) float mel[7][MS];
    NSB_KERNEL_PROLOGUE(TR_STEM)
    const int oh2 = blockIdx.x, b = blockIdx.y, c = threadIdx.x;
    const int M = FULL ? T : PRE_CACHE + 8 * T, t1 = M / 2 + 1, t2 = gridDim.x;
    constexpr int W1 = N_MELS / 2 + 1, W2 = W1 / 2 + 1;                         // 65, 33
    const int slot = FULL ? 0 : slot_of_b[b];
    for (int e = c; e < 7 * MS; e += SUB_CH) {
        const int r = e / MS, m = e % MS - 4, f = 4 * oh2 - 6 + r;              // mel rows 4 oh2 - 6 .. 4 oh2, columns -4 .. 131 (zero outside the image)
        float v = 0.0f;
        if (f >= 0 && f < M && m >= 0 && m < N_MELS) v = FULL ? hist[((size_t)b * M + f) * N_MELS + m] : chunk_mel(hist, mel_new, slot, b, T, f, m);
        mel[r][m + 4] = v;
    }
    float w0[9], w2[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) { w0[k] = w0_t[k * SUB_CH + c]; w2[k] = w2_t[k * SUB_CH + c]; }
    const float bias0 = b0[c], bias2 = b2[c];
    __syncthreads();
    bool hvalid[3];                                                             // conv0 rows 2 oh2 - 2 + kh2 inside the conv0 image (else: the depthwise conv's zero padding)
#pragma unroll
    for (int kh2 = 0; kh2 < 3; ++kh2) { const int h = 2 * oh2 - 2 + kh2; hvalid[kh2] = h >= 0 && h < t1; }
    float win[3][3];
#pragma unroll
    for (int kh2 = 0; kh2 < 3; ++kh2) win[kh2][2] = 0.0f;                        // conv0 column -2: out of range
    for (int ow2 = 0; ow2 < W2; ++ow2) {
        float m[7][5];                                                          // mel rows x columns 4 ow2 - 4 .. 4 ow2
#pragma unroll
        for (int r = 0; r < 7; ++r) {
            const float4 v = *reinterpret_cast<const float4*>(&mel[r][4 * ow2]);
            m[r][0] = v.x; m[r][1] = v.y; m[r][2] = v.z; m[r][3] = v.w; m[r][4] = mel[r][4 * ow2 + 4];
        }
#pragma unroll
        for (int kh2 = 0; kh2 < 3; ++kh2) {
            // conv0 at columns w = 2 ow2 - 1 (mel columns 4 ow2 - 4 ..) and w = 2 ow2 (mel columns 4 ow2 - 2 ..); rows 2 kh2 + kh.
            // Same tap order as the stand-alone conv0 kernel; out-of-image taps contribute fma(0, w, acc) = acc there too.
            float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    a0 = fmaf(m[2 * kh2 + kh][kw], w0[kh * 3 + kw], a0);
                    a1 = fmaf(m[2 * kh2 + kh][kw + 2], w0[kh * 3 + kw], a1);
                }
            }
            win[kh2][0] = win[kh2][2];
            win[kh2][1] = (hvalid[kh2] && ow2 >= 1) ? fmaxf(a0 + bias0, 0.0f) : 0.0f;
            win[kh2][2] = hvalid[kh2] ? fmaxf(a1 + bias0, 0.0f) : 0.0f;
        }
        float acc = 0.0f;
#pragma unroll
        for (int kh2 = 0; kh2 < 3; ++kh2)
#pragma unroll
            for (int kw2 = 0; kw2 < 3; ++kw2) acc = fmaf(win[kh2][kw2], w2[kh2 * 3 + kw2], acc);
        store_pixel(out, ((size_t)b * t2 + oh2) * W2 + ow2, c, acc + bias2, split);
    }
    NSB_KERNEL_EPILOGUE();
}
void launch_stem_conv0_dw(const float* mel_hist, const float* mel_new, const int* slot_of_b, int B, int T, const float* w0_t,
                          const float* b0, const float* w2_t, const float* b2, float* out, cudaStream_t st, int split) {
    const int M = PRE_CACHE + 8 * T, t1 = M / 2 + 1, t2 = t1 / 2 + 1;
    launch_k(stem_conv0_dw_kernel<false>, dim3(t2, B), dim3(SUB_CH), 0, st, mel_hist, mel_new, slot_of_b, T, w0_t, b0, w2_t, b2, out, split);
}
void launch_stem_conv0_dw_full(const float* mel, int B, int M, const float* w0_t, const float* b0, const float* w2_t, const float* b2,
                               float* out, cudaStream_t st, int split) {
    const int t1 = M / 2 + 1, t2 = t1 / 2 + 1;
    launch_k(stem_conv0_dw_kernel<true>, dim3(t2, B), dim3(SUB_CH), 0, st, mel, (const float*)nullptr, (const int*)nullptr, M, w0_t, b0, w2_t, b2, out, split);
}

__global__ void __launch_bounds__(N_MELS) mel_hist_update_kernel(float* __restrict__ hist, const float* __restrict__ mel_new,
                                                                 const int* __restrict__ slot_of_b, int T) {
    NSB_KERNEL_PROLOGUE(TR_MELHIST)
    const int b = blockIdx.x, m = threadIdx.x, slot = slot_of_b[b];
    float v[PRE_CACHE];
#pragma unroll
    for (int f = 0; f < PRE_CACHE; ++f) v[f] = chunk_mel(hist, mel_new, slot, b, T, 8 * T + f, m);   // last 9 of [hist || new]
#pragma unroll
    for (int f = 0; f < PRE_CACHE; ++f) hist[((size_t)slot * PRE_CACHE + f) * N_MELS + m] = v[f];
    NSB_KERNEL_EPILOGUE();
}
void launch_mel_hist_update(float* mel_hist, const float* mel_new, const int* slot_of_b, int B, int T, cudaStream_t st) {
    launch_k(mel_hist_update_kernel, dim3(B), dim3(N_MELS), 0, st, mel_hist, mel_new, slot_of_b, T);
}

__global__ void __launch_bounds__(N_MELS) mel_gather_kernel(const float* __restrict__ hist, const float* __restrict__ mel_new,
                                                            const int* __restrict__ slot_of_b, int T, float* __restrict__ out) {
    NSB_KERNEL_PROLOGUE(TR_OTHER)
    const int f = blockIdx.x, b = blockIdx.y, m = threadIdx.x, M = PRE_CACHE + 8 * T;
    out[((size_t)b * M + f) * N_MELS + m] = chunk_mel(hist, mel_new, slot_of_b[b], b, T, f, m);
    NSB_KERNEL_EPILOGUE();
}
void launch_mel_gather(const float* mel_hist, const float* mel_new, const int* slot_of_b, int B, int T, float* out, cudaStream_t st) {
    launch_k(mel_gather_kernel, dim3(PRE_CACHE + 8 * T, B), dim3(N_MELS), 0, st, mel_hist, mel_new, slot_of_b, T, out);
}

// depthwise 3x3 stride 2 (+bias, no activation), NHWC with C = 256: thread = channel, one CTA = one output ROW (b, oh): a 3x3
// register window slides along the row (6 new loads per output instead of 9, and 17x fewer CTAs than one per output pixel:
// at 256 streams that was 39 168 one-pixel CTAs). Taps outside the image are skipped, in the (kh, kw) order of the reference loop.
__global__ void __launch_bounds__(SUB_CH) dwconv_s2_kernel(const float* __restrict__ in, int H, int W, const float* __restrict__ w_t,
                                                           const float* __restrict__ bias, float* __restrict__ out, int split) {
    NSB_KERNEL_PROLOGUE(TR_DWCONV)
    const int oh = blockIdx.x, b = blockIdx.y, c = threadIdx.x;
    const int Ho = gridDim.x, Wo = W / 2 + 1;
    float w[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) w[k] = w_t[k * SUB_CH + c];
    const float bs = bias[c];
    const float* row[3]; bool rv[3];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
        const int ih = 2 * oh + kh - 2;
        rv[kh] = ih >= 0 && ih < H;
        row[kh] = in + (((size_t)b * H + (rv[kh] ? ih : 0)) * W) * SUB_CH + c;
    }
    float win[3][3];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) win[kh][2] = 0.0f;                            // column -2: outside (never used: its tap is skipped)
#pragma unroll 4
    for (int ow = 0; ow < Wo; ++ow) {
        const int i1 = 2 * ow - 1, i2 = 2 * ow;                                  // columns 2 ow - 2 (carried), 2 ow - 1, 2 ow
        const bool v0 = ow >= 1, v1 = i1 >= 0 && i1 < W, v2 = i2 < W;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            win[kh][0] = win[kh][2];
            win[kh][1] = (rv[kh] && v1) ? row[kh][(size_t)i1 * SUB_CH] : 0.0f;
            win[kh][2] = (rv[kh] && v2) ? row[kh][(size_t)i2 * SUB_CH] : 0.0f;
        }
        float acc = 0.0f;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            if (!rv[kh]) continue;
            if (v0) acc = fmaf(win[kh][0], w[kh * 3 + 0], acc);
            if (v1) acc = fmaf(win[kh][1], w[kh * 3 + 1], acc);
            if (v2) acc = fmaf(win[kh][2], w[kh * 3 + 2], acc);
        }
        store_pixel(out, ((size_t)b * Ho + oh) * Wo + ow, c, acc + bs, split);
    }
    NSB_KERNEL_EPILOGUE();
}
void launch_dwconv_s2(const float* in, int B, int H, int W, const float* w_t, const float* bias, float* out, cudaStream_t st, int split) {
    launch_k(dwconv_s2_kernel, dim3(H / 2 + 1, B), dim3(SUB_CH), 0, st, in, H, W, w_t, bias, out, split);
}

__global__ void __launch_bounds__(256) split_tf32_kernel(const float4* __restrict__ in, float* __restrict__ out, size_t n_vec, int K) {
    NSB_KERNEL_PROLOGUE(TR_OTHER)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = in[i];
        const size_t e = i * 4, row = e / (size_t)K, col = e % (size_t)K;
        float4 hi, lo;
        split_tf32(v.x, hi.x, lo.x); split_tf32(v.y, hi.y, lo.y); split_tf32(v.z, hi.z, lo.z); split_tf32(v.w, hi.w, lo.w);
        float* o = out + row * 2 * (size_t)K + col;
        *reinterpret_cast<float4*>(o) = hi; *reinterpret_cast<float4*>(o + K) = lo;
    }
    NSB_KERNEL_EPILOGUE();
}
void launch_split_tf32(const float* in, float* out, size_t rows, int K, cudaStream_t st) {
    if (K % 4 != 0) throw CudaError("split_tf32: K must be a multiple of 4");
    const size_t n_vec = rows * (size_t)K / 4;
    if (!n_vec) return;
    const int blocks = (int)std::min<size_t>((n_vec + 255) / 256, 148 * 8);
    launch_k(split_tf32_kernel, dim3(blocks), dim3(256), 0, st, (const float4*)in, out, n_vec, K);
}

}  // namespace nsb
