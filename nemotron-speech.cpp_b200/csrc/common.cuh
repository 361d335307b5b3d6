// common.cuh -- shared declarations for the sm_100a engine (internal; the public face is include/nsb200.h)
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <utility>

namespace nsb {

// ------------------------------------------------------------------------------------------
// model constants (nemotron-speech-streaming-en-0.6b; reference src/nemo-ggml.h:37-49, :129-133)
// ------------------------------------------------------------------------------------------
constexpr int D_MODEL = 1024, D_FF = 4096, N_HEADS = 8, D_HEAD = 128;
constexpr int N_MELS = 128, N_FFT = 512, N_BINS = 257, HOP = 160, WIN = 400;
constexpr int ATT_L = 70;            // att_left_context  (nemo-stream.h:25)
constexpr int CONV_K = 9;            // conv_kernel_size  (nemo-stream.h:30)
constexpr int PRE_CACHE = 9;         // pre_encode_cache_size (nemo-stream.h:57)
constexpr int DROP_PRE = 2;          // drop_extra_pre_encoded (nemo-stream.h:55)
constexpr int SUB_CH = 256, SUB_W = 17;
constexpr int HID = 640, JOINT = 640, VOCAB = 1025, BLANK = 1024;
constexpr int MAX_SYMBOLS = 10;      // nemo-stream.cpp:797

struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };

#define NSB_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            throw ::nsb::CudaError(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + \
                                   std::to_string(__LINE__) + ")");                                 \
    } while (0)

// ------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL). Every kernel of a step is launched with the programmatic-stream-serialization
// attribute: its CTAs may become resident (and run their input-independent prologue: barrier init, TMEM allocation,
// tensor-map prefetch, weight-tile TMA) while the previous kernel drains. pdl_wait() blocks until the previous grid has
// completed and its writes are visible -- it must precede the first access to anything a predecessor produces or still
// reads; pdl_trigger() lets the NEXT grid start launching. Without the attribute both are no-ops.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pdl_enabled();            // engine.cu (NSB_NO_PDL=1 disables)
// The first kernel behind a cross-stream event wait is launched WITHOUT the programmatic attribute (a plain, full dependency on
// everything before it): pdl_skip_next() arms a one-shot, per-thread switch that the next launch_k / launch_k_cluster consumes.
inline bool& pdl_skip_flag() { static thread_local bool f = false; return f; }
inline void pdl_skip_next() { pdl_skip_flag() = true; }
inline bool pdl_take() { bool& f = pdl_skip_flag(); const bool on = pdl_enabled() && !f; f = false; return on; }

template <typename... P, typename... A>
inline void launch_k(void (*kern)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_take() ? 1 : 0;
    NSB_CUDA(cudaLaunchKernelEx(&cfg, kern, P(std::forward<A>(args))...));
}

// same, with a thread-block cluster of `cluster_x` CTAs along x
template <typename... P, typename... A>
inline void launch_k_cluster(void (*kern)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x, A&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[2]; int n = 0;
    if (cluster_x > 1) { at[n].id = cudaLaunchAttributeClusterDimension; at[n].val.clusterDim.x = cluster_x; at[n].val.clusterDim.y = 1; at[n].val.clusterDim.z = 1; ++n; }
    if (pdl_take()) { at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[n].val.programmaticStreamSerializationAllowed = 1; ++n; }
    cfg.attrs = at; cfg.numAttrs = n;
    NSB_CUDA(cudaLaunchKernelEx(&cfg, kern, P(std::forward<A>(args))...));
}

// cudaFuncSetAttribute acts on the CURRENT device only. One process may hold one engine per GPU (one host thread each), so the
// "already configured" memo of a kernel is kept per device ordinal, never per process: `slots` is a function-local
// `static std::atomic<size_t> [MAX_DEVICES]` (zero-initialised), raised to the largest dynamic shared-memory size asked for so far.
constexpr int MAX_DEVICES = 64;
template <typename K>
inline void ensure_dyn_smem(K kern, size_t bytes, std::atomic<size_t>* slots, bool max_carveout = false) {
    int dev = 0;
    NSB_CUDA(cudaGetDevice(&dev));
    std::atomic<size_t>& done = slots[dev & (MAX_DEVICES - 1)];
    if (bytes <= done.load(std::memory_order_acquire)) return;
    NSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    if (max_carveout) NSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    done.store(bytes, std::memory_order_release);
}

// ------------------------------------------------------------------------------------------
// In-situ device tracing (nsb_trace_enable): thread 0 of block (0,0,0) of every kernel claims a record and stamps
// %globaltimer at a few points. Unlike ncu (which serialises launches) this shows the timeline INSIDE the CUDA graph:
// launch gaps, PDL overlap, how long a kernel sat in griddepcontrol.wait. Off = one constant-cache load per kernel.
//   t[0] block 0 started   t[1] griddepcontrol.wait returned   t[2] block 0 finished
//   GEMM: t[1] prologue done, t[2] wait returned (epilogue warp), t[3] accumulator complete, t[4] epilogue done
// ------------------------------------------------------------------------------------------
struct TraceRec { unsigned long long t[6]; int tag; int grid; };
struct TraceBuf { unsigned n; unsigned cap; unsigned pad[2]; TraceRec rec[1]; };
enum TraceTag : int { TR_LOGMEL = 1, TR_STEM, TR_DWCONV, TR_MELHIST, TR_GEMM_SIMT, TR_GEMM_TC, TR_GEMM_Q8, TR_LN, TR_LN2, TR_ATTN, TR_CONVMOD, TR_ADVANCE, TR_DECODE, TR_OTHER };
static __constant__ TraceBuf* c_trace;          // one copy per translation unit, bound by NSB_DEFINE_TRACE_BINDER
#define NSB_DEFINE_TRACE_BINDER(name) \
    void name(TraceBuf* p) { NSB_CUDA(cudaMemcpyToSymbol(c_trace, &p, sizeof(p))); }
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ bool trace_thread() { return threadIdx.x == 0 && (blockIdx.x | blockIdx.y | blockIdx.z) == 0; }
__device__ __forceinline__ int trace_begin(int tag) {
    TraceBuf* tb = c_trace;
    if (!tb) return -1;
    const unsigned s = atomicAdd(&tb->n, 1u);
    if (s >= tb->cap) return -1;
    tb->rec[s].tag = tag; tb->rec[s].grid = (int)(gridDim.x * gridDim.y * gridDim.z); tb->rec[s].t[0] = gtime();
    return (int)s;
}
__device__ __forceinline__ void trace_mark(int slot, int i) { if (slot >= 0) c_trace->rec[slot].t[i] = gtime(); }
// The common prologue of a simple kernel, in two halves:
//   NSB_KERNEL_BEGIN  trace start
//   ... input-independent loads of this kernel go here (weights, K/V ring rows, conv state): they overlap the predecessor
//       whenever this grid was launched early ...
//   NSB_KERNEL_WAIT   griddepcontrol.wait (predecessor COMPLETE and flushed, hence transitively everything before it), then
//                     launch_dependents. Measured (profiles/r01_notes.md): triggering at kernel ENTRY instead lets the whole
//                     layer pile up on the SMs ~15 us ahead, which bought nothing and lengthened the completion -> release
//                     gaps (19 -> 25 us per layer); triggering after the wait keeps the launch front one or two kernels ahead,
//                     enough to hide the ~2 us launch latency behind this kernel's body. The GEMMs trigger right after their
//                     prologue (they have real pre-wait work to overlap: the weight-slab prefetch).
#define NSB_KERNEL_BEGIN(tag)                                      \
    int tr_slot = -1;                                              \
    if (trace_thread()) tr_slot = trace_begin(tag);
#define NSB_KERNEL_WAIT()                                          \
    pdl_wait(); pdl_trigger();                                     \
    if (tr_slot >= 0) trace_mark(tr_slot, 1);
#define NSB_KERNEL_PROLOGUE(tag) NSB_KERNEL_BEGIN(tag) NSB_KERNEL_WAIT()
#define NSB_KERNEL_EPILOGUE() do { if (tr_slot >= 0) trace_mark(tr_slot, 2); } while (0)

// Epilogues shared by the SIMT and the tcgen05 GEMM
enum Epi : int {
    EPI_NONE = 0,      // C = acc (+bias)
    EPI_RELU = 1,      // C = relu(acc + bias)
    EPI_SILU = 2,      // C = silu(acc)
    EPI_RESID = 3,     // C (f32, in place) += alpha * acc
    EPI_PARTIAL = 4    // split-K: C[z][M][N] (f32 workspace) = partial acc of k-slice z; reduced by the following LayerNorm kernel
};
enum OutType : int { OUT_F32 = 0, OUT_F16 = 1, OUT_BF16 = 2 };

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
__device__ __forceinline__ float silu_exact(float x) { return x / (1.0f + expf(-x)); }
__device__ __forceinline__ float sigmoid_exact(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }

// store one value as OutType ot at element index i of a raw pointer
__device__ __forceinline__ void store_out(void* p, size_t i, float v, int ot) {
    if (ot == OUT_F32) ((float*)p)[i] = v;
    else if (ot == OUT_F16) ((__half*)p)[i] = __float2half_rn(v);
    else ((__nv_bfloat16*)p)[i] = __float2bfloat16_rn(v);
}
__device__ __forceinline__ float load_kv(const void* p, size_t i, int kv) {
    if (kv == 0) return ((const float*)p)[i];
    if (kv == 1) return __half2float(((const __half*)p)[i]);
    return __bfloat162float(((const __nv_bfloat16*)p)[i]);
}
__device__ __forceinline__ float round_kv(float v, int kv) {
    if (kv == 0) return v;
    if (kv == 1) return __half2float(__float2half_rn(v));
    return __bfloat162float(__float2bfloat16_rn(v));
}
__device__ __forceinline__ void store_kv(void* p, size_t i, float v, int kv) {
    if (kv == 0) ((float*)p)[i] = v;
    else if (kv == 1) ((__half*)p)[i] = __float2half_rn(v);
    else ((__nv_bfloat16*)p)[i] = __float2bfloat16_rn(v);
}
// 3xTF32: v = hi + lo with hi, lo exactly representable in tf32 (10 explicit mantissa bits); the dropped remainder is <= 2^-22 |v|.
// a . w ~= a_hi w_hi + a_hi w_lo + a_lo w_hi on kind::tf32 tensor cores with fp32 accumulation: fp32-class accuracy for the
// matrices every GGUF keeps in F32 (subsampling stem, joint.enc), which the reference multiplies in fp32.
__device__ __forceinline__ float to_tf32(float v) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v)); return __uint_as_float(r); }
__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) { hi = to_tf32(v); lo = to_tf32(v - hi); }
// NHWC pixel store, C = 256: plain, or [hi (256) | lo (256)] for a 3xTF32 consumer
__device__ __forceinline__ void store_pixel(float* out, size_t pix, int c, float v, int split) {
    if (!split) { out[pix * 256 + c] = v; return; }
    float hi, lo; split_tf32(v, hi, lo);
    out[pix * 512 + c] = hi; out[pix * 512 + 256 + c] = lo;
}
inline size_t kv_elem_size(int kv) { return kv == 0 ? 4 : 2; }
inline size_t out_elem_size(int ot) { return ot == OUT_F32 ? 4 : 2; }

}  // namespace nsb
