// C[M,N] = A[M,K] W[N,K]^T with fused epilogues — SIMT, true fp32 FMA for every dtype.
// This is the fp32 verification path (rtol 1e-4 rules out TF32) and the any-shape fallback
// of pg_gemm; bf16/f16 GEMMs on the prefill / vision path go to the tcgen05 kernel
// (gemm_tcgen05.cu) when the shape allows.
#include "common.cuh"

namespace pg {

constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16, SG_THREADS = 256, SG_PAD = 4;

template <typename T>
__device__ __forceinline__ void sg_load_tile(float (*s)[SG_BM + SG_PAD], const T* __restrict__ base, int ld,
                                             int row0, int rows, int k0, int K) {
  constexpr int V = Vec<T>::N;
  constexpr int VPR = SG_BK / V;  // vectors per tile row
  for (int e = threadIdx.x; e < SG_BM * VPR; e += SG_THREADS) {
    const int r = e / VPR, kv = (e % VPR) * V;
    float f[V];
    if (row0 + r < rows && k0 + kv < K) unpack<T>(ldg_cached(base + (size_t)(row0 + r) * ld + k0 + kv), f);
    else
#pragma unroll
      for (int i = 0; i < V; ++i) f[i] = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) s[kv + i][r] = f[i];
  }
}

template <typename T, int EPI>
__global__ void __launch_bounds__(SG_THREADS)
gemm_simt_kernel(void* __restrict__ Cv, const T* __restrict__ A, const T* __restrict__ W,
                 const T* __restrict__ bias, const T* __restrict__ R, int M, int N, int K, int lda, int ldw,
                 int ldc, int ldr, int res_mod, int out_f32) {
  constexpr bool DUAL = (EPI == PG_EPI_GEGLU);
  __shared__ __align__(16) float As[SG_BK][SG_BM + SG_PAD];
  __shared__ __align__(16) float Bs[SG_BK][SG_BN + SG_PAD];
  __shared__ __align__(16) float Us[DUAL ? SG_BK : 1][SG_BN + SG_PAD];
  const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  float acc[4][4], acu[DUAL ? 4 : 1][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j] = 0.f; if (DUAL) acu[i][j] = 0.f; }

  for (int k0 = 0; k0 < K; k0 += SG_BK) {
    sg_load_tile<T>(As, A, lda, m0, M, k0, K);
    sg_load_tile<T>(Bs, W, ldw, n0, N, k0, K);
    if (DUAL) sg_load_tile<T>(Us, W + (size_t)N * ldw, ldw, n0, N, k0, K);
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SG_BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      if (DUAL) {
        const float4 u = *reinterpret_cast<const float4*>(&Us[kk][tx * 4]);
        const float uv[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acu[i][j] = fmaf(av[i], uv[j], acu[i][j]);
      }
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (EPI == PG_EPI_BIAS || EPI == PG_EPI_BIAS_GELU || EPI == PG_EPI_BIAS_RES) v += to_f<T>(bias[n]);
      v = rnd<T>(v);
      if (EPI == PG_EPI_BIAS_GELU) v = rnd<T>(gelu_tanh(v));
      if (EPI == PG_EPI_GEGLU) v = rnd<T>(rnd<T>(gelu_tanh(v)) * rnd<T>(acu[DUAL ? i : 0][j]));
      if (EPI == PG_EPI_BIAS_RES || EPI == PG_EPI_RES) {
        const int rm = res_mod > 0 ? (m % res_mod) : m;
        v = rnd<T>(v + to_f<T>(R[(size_t)rm * ldr + n]));
      }
      if (out_f32) reinterpret_cast<float*>(Cv)[(size_t)m * ldc + n] = v;
      else reinterpret_cast<T*>(Cv)[(size_t)m * ldc + n] = from_f<T>(v);
    }
  }
}

template <typename T>
static int launch_simt(void* C, const void* A, const void* W, const void* bias, const void* R, int M, int N,
                       int K, int lda, int ldw, int ldc, int ldr, int res_mod, int epi, int out_f32,
                       cudaStream_t st) {
  dim3 grid(cdiv(N, SG_BN), cdiv(M, SG_BM));
#define PG_SG(E)                                                                                      \
  gemm_simt_kernel<T, E><<<grid, SG_THREADS, 0, st>>>(C, (const T*)A, (const T*)W, (const T*)bias,   \
                                                      (const T*)R, M, N, K, lda, ldw, ldc, ldr, res_mod, out_f32)
  switch (epi) {
    case PG_EPI_NONE: PG_SG(PG_EPI_NONE); break;
    case PG_EPI_BIAS: PG_SG(PG_EPI_BIAS); break;
    case PG_EPI_BIAS_GELU: PG_SG(PG_EPI_BIAS_GELU); break;
    case PG_EPI_BIAS_RES: PG_SG(PG_EPI_BIAS_RES); break;
    case PG_EPI_RES: PG_SG(PG_EPI_RES); break;
    case PG_EPI_GEGLU: PG_SG(PG_EPI_GEGLU); break;
    default: set_error("gemm: bad epilogue %d", epi); return PG_ERR_INVALID;
  }
#undef PG_SG
  return check_launch("gemm_simt");
}

int gemm_simt(void* C, const void* A, const void* W, const void* bias, const void* R, int M, int N, int K,
              int lda, int ldw, int ldc, int ldr, int res_mod, int epi, int out_f32, int dtype,
              cudaStream_t st) {
  PG_DISPATCH_DTYPE(dtype, T, {
    constexpr int V = Vec<T>::N;
    PG_REQUIRE(K % V == 0 && lda % V == 0 && ldw % V == 0, "gemm: K=%d lda=%d ldw=%d must be multiples of %d",
               K, lda, ldw, V);
    return launch_simt<T>(C, A, W, bias, R, M, N, K, lda, ldw, ldc, ldr, res_mod, epi, out_f32, st);
  });
  return PG_OK;
}

}  // namespace pg
