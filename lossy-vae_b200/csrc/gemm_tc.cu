// tcgen05 tensor-core GEMM (placeholder until the TMA/UMMA kernel lands in this file).
#include "common.cuh"
namespace lvae {
int64_t gemm_tc_workspace_bytes(const lvae_gemm_desc*) { return 0; }
int gemm_tc_launch(const lvae_gemm_desc*, cudaStream_t) {
  set_error("tensor-core GEMM not built into this library yet");
  return LVAE_E_UNSUPPORTED;
}
}
