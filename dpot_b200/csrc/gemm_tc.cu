// tcgen05 / TMEM / TMA 3xTF32 GEMM engine (DPOT_GEMM_TC).  Placeholder until the engine lands.
#include "common.cuh"
#include "gemm_common.cuh"

namespace dpot {
bool gemm_tc_supports(const GemmDev&, int) { return false; }
int gemm_tc_launch(const GemmDev&, int, cudaStream_t) {
  set_error("tcgen05 engine not built");
  return DPOT_E_UNSUPPORTED;
}
}  // namespace dpot

extern "C" int dpot_tc_available(void) { return 0; }
