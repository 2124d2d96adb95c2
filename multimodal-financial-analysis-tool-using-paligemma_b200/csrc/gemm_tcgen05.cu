// tcgen05 / TMEM / TMA GEMM — placeholder until the kernel lands; pg_gemm routes to SIMT.
#include "common.cuh"
namespace pg {
bool gemm_tc_supported(int, int, int, int, int, int, int, int, int) { return false; }
int gemm_tc(void*, const void*, const void*, const void*, const void*, int, int, int, int, int, int, int, int,
            int, int, int, cudaStream_t) {
  set_error("tcgen05 GEMM not built");
  return PG_ERR_INVALID;
}
}  // namespace pg
