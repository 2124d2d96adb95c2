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
bool gemm_tc_wide_supported(int M, int N, int K, int epi);
int gemm_tc_wide(void* C, const void* A, const void* W, const void* bias, const void* R, int M, int N, int K, int lda,
                 int ldw, int ldc, int ldr, int epi, int out_f32, int dtype, cudaStream_t st);
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
  if (impl == 3) {  // experimental swap-AB kernel for prompt-sized row counts (gemm_tcgen05_wide.cu)
    PG_REQUIRE(tc_ok && res_mod == 0 && gemm_tc_wide_supported(M, N, K, epilogue),
               "gemm: the wide swap-AB path does not support this problem (M=%d N=%d K=%d)", M, N, K);
    return gemm_tc_wide(C, A, W, bias, R, M, N, K, lda, ldw, ldc, ldr, epilogue, out_f32, dtype, st);
  }
  if (impl == 2) {
    PG_REQUIRE(tc_ok, "gemm: tcgen05 path does not support this problem (M=%d N=%d K=%d dtype=%d)", M, N, K, dtype);
    return gemm_tc(C, A, W, bias, R, M, N, K, lda, ldw, ldc, ldr, res_mod, epilogue, out_f32, dtype, st);
  }
  if (impl == 0 && tc_ok)
    return gemm_tc(C, A, W, bias, R, M, N, K, lda, ldw, ldc, ldr, res_mod, epilogue, out_f32, dtype, st);
  return gemm_simt(C, A, W, bias, R, M, N, K, lda, ldw, ldc, ldr, res_mod, epilogue, out_f32, dtype, st);
}
