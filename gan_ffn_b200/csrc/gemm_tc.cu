// tcgen05 (kind::tf32) 3xTF32 GEMM engine -- placeholder until the tensor path lands.
#include "kernels.h"
namespace ganffn {
bool gemm_tc_supported(bool, bool, int, int, int, int, int, int, const void*, const void*) { return false; }
int64_t gemm_tc_scratch_floats(int, int, int) { return 0; }
int gemm_tc(const float*, int, bool, const float*, int, bool, float*, int, int, int, int, const Epilogue&, float*, int64_t,
            cudaStream_t) {
  set_error("gemm_tc: not built");
  return GANFFN_ERR_ARG;
}
}  // namespace ganffn
