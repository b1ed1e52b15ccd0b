// tcgen05 / TMEM / TMA GEMM kernels (placeholder until the tensor-core engine lands).
#include "common.cuh"
namespace pcadv {
int tc_linear(const pcadv_linear_args&, cudaStream_t) {
  set_error("tensor-core engine not built in this library");
  return 10;
}
int tc_wgrad(const pcadv_wgrad_args&, cudaStream_t) {
  set_error("tensor-core engine not built in this library");
  return 10;
}
}  // namespace pcadv
