// tcgen05 3xTF32 GEMM (placeholder until the tensor-core kernel lands: reports "unsupported").
#include "gemm.cuh"

namespace cgc {
bool gemm_tc_supported(const GemmParams&, int, int) { return false; }
int gemm_tc(GemmParams, int, int, void*, int64_t, cudaStream_t) {
  set_error("tcgen05 GEMM not built");
  return CGC_ERR_UNSUPPORTED;
}
}  // namespace cgc
