// tcgen05 TF32 implicit-GEMM engine (placeholder until the kernel lands).
#include "common.cuh"
namespace frido {
int conv2d_tc(const FridoConvParams* p, cudaStream_t s) {
  (void)p; (void)s;
  return set_error(FRIDO_E_ARG, "conv2d: tcgen05 engine not built");
}
}  // namespace frido
