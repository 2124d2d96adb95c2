// pg_gemm: front door of every projection on the prefill / vision path.
#include "common.cuh"

namespace pg {
int gemm_simt(void* C, const void* A, const void* W, const void* bias, const void* R, int M, int N, int K,
              int lda, int ldw, int ldc, int ldr, int res_mod, int epi, int out_f32, int dtype,
              cudaStream_t st);
int gemm_tc(void* C, const void* A, const void* W, const void* bias, const void* R, int M, int N, int K,
            int lda, int ldw, int ldc, int ldr, int res_mod, int epi, int out_f32, int dtype,
            cudaStream_t st);
bool gemm_tc_supported(int M, int N, int K, int lda, int ldw, int ldc, int epi, int out_f32, int dtype);
bool gemm_tc_swap_wanted(int M, int N, int K, int epi, int out_f32, int res_mod);
void gemm_tc_swap_force(bool on);
}  // namespace pg

using namespace pg;

extern "C" int pg_gemm(void* C, const void* A, const void* W, const void* bias, const void* R, int M, int N,
                       int K, int lda, int ldw, int ldc, int ldr, int res_mod, int epilogue, int out_f32,
                       int impl, int dtype, void* stream) {
  if (M <= 0 || N <= 0) return PG_OK;
  PG_REQUIRE(K > 0 && lda >= K && ldw >= K && ldc >= N, "gemm: bad leading dimensions");
  const bool need_bias = epilogue == PG_EPI_BIAS || epilogue == PG_EPI_BIAS_GELU || epilogue == PG_EPI_BIAS_RES;
  const bool need_res = epilogue == PG_EPI_BIAS_RES || epilogue == PG_EPI_RES;
  PG_REQUIRE(!need_bias || bias, "gemm: epilogue %d needs a bias", epilogue);
  PG_REQUIRE(!need_res || (R && ldr >= N), "gemm: epilogue %d needs a residual", epilogue);
  cudaStream_t st = (cudaStream_t)stream;
  const bool tc_ok = gemm_tc_supported(M, N, K, lda, ldw, ldc, epilogue, out_f32, dtype);
  if (impl == 3) {  // the CTA-pair swap-AB kernel for prompt-sized row counts, whatever the per-projection default says
    gemm_tc_swap_force(true);
    const bool ok = tc_ok && gemm_tc_swap_wanted(M, N, K, epilogue, out_f32, res_mod);
    int rc = PG_ERR_INVALID;
    if (ok) rc = gemm_tc(C, A, W, bias, R, M, N, K, lda, ldw, ldc, ldr, res_mod, epilogue, out_f32, dtype, st);
    else set_error("gemm: the pair swap-AB kernel does not take this problem (M=%d N=%d K=%d epilogue %d; workspace set?)", M, N, K, epilogue);
    gemm_tc_swap_force(false);
    return rc;
  }
  if (impl == 2) {
    PG_REQUIRE(tc_ok, "gemm: tcgen05 path does not support this problem (M=%d N=%d K=%d dtype=%d)", M, N, K, dtype);
    return gemm_tc(C, A, W, bias, R, M, N, K, lda, ldw, ldc, ldr, res_mod, epilogue, out_f32, dtype, st);
  }
  if (impl == 0 && tc_ok)
    return gemm_tc(C, A, W, bias, R, M, N, K, lda, ldw, ldc, ldr, res_mod, epilogue, out_f32, dtype, st);
  return gemm_simt(C, A, W, bias, R, M, N, K, lda, ldw, ldc, ldr, res_mod, epilogue, out_f32, dtype, st);
}
