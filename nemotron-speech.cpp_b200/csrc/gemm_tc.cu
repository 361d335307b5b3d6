// placeholder until the tcgen05 kernel lands (next commit)
#include "kernels.cuh"
namespace nsb {
void launch_gemm_tc(const GemmArgs&, int, cudaStream_t) { throw CudaError("tcgen05 GEMM not built yet: use NSB_COMPUTE_F32"); }
}
