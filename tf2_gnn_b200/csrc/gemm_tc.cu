// 3xTF32 tcgen05 node-level GEMM (placeholder until the tensor-core path lands).
#include "gemm.cuh"
namespace tfgnn {
bool gemm_tc_supported(long long, int, int, const float*, int, const float*, int) { return false; }
size_t gemm_tc_packed_bytes(int, int) { return 0; }
int launch_pack_weights_tc(const float*, int, int, int, float*, cudaStream_t) {
  set_error(TFGNN_ERR_UNSUPPORTED, "tcgen05 GEMM not built");
  return TFGNN_ERR_UNSUPPORTED;
}
int launch_gemm_tc(const float*, int, const float*, float*, int, long long, int, int, const GemmEpilogue&, cudaStream_t) {
  set_error(TFGNN_ERR_UNSUPPORTED, "tcgen05 GEMM not built");
  return TFGNN_ERR_UNSUPPORTED;
}
}  // namespace tfgnn
