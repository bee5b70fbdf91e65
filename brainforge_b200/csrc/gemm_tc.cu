// Placeholder until the tcgen05 kernel lands: nothing is "supported", so every call takes the SIMT path.
#include "gemm_tc.cuh"
namespace bf {
bool tc_gemm_supported(int, int, int, bool, bool) { return false; }
int tc_gemm(const float*, long, bool, const float*, long, bool, float*, long, int, int, int, const float*, int, int,
            cudaStream_t) {
  return fail(BF_ERR_ARG, "tc_gemm: not built");
}
}  // namespace bf
