// Split-K companion of gemm_tcgen05.cu for GEMMs with few output tiles and a long K (prefill with a few
// hundred tokens: o_proj, down_proj, q/k/v; SigLIP at batch 1).  With 128-row tiles a 260 x 2048 output has
// 48-96 tiles, so most of the 148 SMs would idle while a handful stream all the weights.  Here a cluster
// of S CTAs shares one output tile: CTA r runs the TMA -> tcgen05.mma pipeline over its 1/S of K into its
// own TMEM accumulator, the non-leaders park their fp32 partial tile in shared memory, and after one
// cluster barrier the leader adds them through distributed shared memory (ld.shared::cluster) and
// applies the fused epilogue.  No atomics, no global workspace, deterministic summation order.
#include "tc_common.cuh"

namespace pg {
namespace tc {

constexpr int SK_BM = 128, SK_BK = 64, SK_STAGES = 6;

struct SkParams {
  void* C;
  const void* bias;
  const void* R;
  int M, N, K, ldc, ldr, res_mod, out_f32;
};

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ float ld_dsmem_f32(uint32_t local_addr, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(rank));
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
  return v;
}

template <typename T, int EPI, int BN, int S>
__global__ void __launch_bounds__(256, 1)
gemm_tc_splitk_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, SkParams p) {
  constexpr int A_BYTES = SK_BM * SK_BK * 2, W_BYTES = BN * SK_BK * 2, STAGE_BYTES = A_BYTES + W_BYTES;
  constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  constexpr uint32_t IDESC = umma_idesc(std::is_same<T, bf16>::value ? 1 : 0, SK_BM, BN);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = smem_base + SK_STAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (SK_STAGES + s); };
  const uint32_t tfull_bar = bars + 8u * (2 * SK_STAGES), tmem_slot = tfull_bar + 8;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  float* part = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)));  // [BN][128] fp32, reuses the stage ring

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  const int tile = blockIdx.x / S;
  const int m_tiles = (p.M + SK_BM - 1) / SK_BM;
  const int n_blk = tile / m_tiles, m_blk = tile % m_tiles;
  const int k_blocks = (p.K + SK_BK - 1) / SK_BK, per = (k_blocks + S - 1) / S;
  const int kb0 = (int)rank * per, kb1 = min(k_blocks, kb0 + per);

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < SK_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tfull_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0 && lane == 0) {
    int stage = 0;
    uint32_t phase = 0;
    pdl_wait();
    for (int kb = kb0; kb < kb1; ++kb) {
      mbar_wait(empty_bar(stage), phase ^ 1);
      const uint32_t sa = smem_base + stage * STAGE_BYTES;
      mbar_expect_tx(full_bar(stage), STAGE_BYTES);
      tma_load_2d(sa, &map_a, full_bar(stage), kb * SK_BK, m_blk * SK_BM);
      tma_load_2d(sa + A_BYTES, &map_w, full_bar(stage), kb * SK_BK, n_blk * BN);
      if (++stage == SK_STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1 && lane == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = kb0; kb < kb1; ++kb) {
      mbar_wait(full_bar(stage), phase);
      tc_fence_after();
      const uint32_t sa = smem_base + stage * STAGE_BYTES;
#pragma unroll
      for (int k = 0; k < SK_BK / 16; ++k)
        umma(tmem_base, umma_desc(sa + k * 32), umma_desc(sa + A_BYTES + k * 32), IDESC, (kb > kb0 || k > 0) ? 1u : 0u);
      umma_commit(empty_bar(stage));
      if (++stage == SK_STAGES) { stage = 0; phase ^= 1; }
    }
    umma_commit(tfull_bar);
  }
  __syncwarp();
  const int q = warp & 3, r = q * 32 + lane;
  const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
  const bool have_k = kb1 > kb0;
  if (warp >= 4) {
    pdl_wait();               // residual reads / output writes only after the predecessor has finished
    mbar_wait(tfull_bar, 0);  // every MMA of this CTA has completed: accumulator final, stage ring idle
    tc_fence_after();
    if (rank != 0) {
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        float v[32];
        tmem_ld32(t_row + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) part[(c * 32 + j) * 128 + r] = have_k ? v[j] : 0.f;
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();  // partial tiles are visible cluster-wide
  if (rank == 0 && warp >= 4) {
    tc_fence_after();
    const int m = m_blk * SK_BM + r;
    T* Ct = reinterpret_cast<T*>(p.C);
    float* Cf = reinterpret_cast<float*>(p.C);
    const T* bias = reinterpret_cast<const T*>(p.bias);
    const T* R = reinterpret_cast<const T*>(p.R);
    const uint32_t part_addr = smem_base;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      float v[32];
      tmem_ld32(t_row + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int peer = 1; peer < S; ++peer)
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += ld_dsmem_f32(part_addr + (uint32_t)(((c * 32 + j) * 128 + r) * 4), peer);
      const int n0 = n_blk * BN + c * 32;
      if (m < p.M && n0 < p.N) {
        const int rm = p.res_mod > 0 ? (m % p.res_mod) : m;
#pragma unroll
        for (int j0 = 0; j0 < 32; j0 += 8) {
          if (n0 + j0 >= p.N) break;
          float o[8], bb[8], rr[8];
          if (EPI == PG_EPI_BIAS || EPI == PG_EPI_BIAS_GELU || EPI == PG_EPI_BIAS_RES) unpack<T>(ldg_cached(bias + n0 + j0), bb);
          if (EPI == PG_EPI_BIAS_RES || EPI == PG_EPI_RES) unpack<T>(ldg_cached(R + (size_t)rm * p.ldr + n0 + j0), rr);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float x = v[j0 + j];
            if (EPI == PG_EPI_BIAS || EPI == PG_EPI_BIAS_GELU || EPI == PG_EPI_BIAS_RES) x += bb[j];
            x = rnd<T>(x);
            if (EPI == PG_EPI_BIAS_GELU) x = rnd<T>(gelu_tanh_fast(x));
            if (EPI == PG_EPI_BIAS_RES || EPI == PG_EPI_RES) x = rnd<T>(x + rr[j]);
            o[j] = x;
          }
          if (p.out_f32) {
            float4* dst = reinterpret_cast<float4*>(Cf + (size_t)m * p.ldc + n0 + j0);
            dst[0] = make_float4(o[0], o[1], o[2], o[3]);
            dst[1] = make_float4(o[4], o[5], o[6], o[7]);
          } else {
            *reinterpret_cast<uint4*>(Ct + (size_t)m * p.ldc + n0 + j0) = pack<T>(o);
          }
        }
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();  // nobody leaves while the leader may still read its shared memory
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

template <typename T, int EPI, int BN, int S>
static int launch_sk(const CUtensorMap& ma, const CUtensorMap& mw, const SkParams& p, cudaStream_t st) {
  const size_t smem = 1024 + (size_t)SK_STAGES * (SK_BM * SK_BK * 2 + BN * SK_BK * 2) + 8 * (2 * SK_STAGES + 2) + 16;
  auto kern = gemm_tc_splitk_kernel<T, EPI, BN, S>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    set_error("gemm_tc_splitk: cannot reserve %zu B of shared memory", smem);
    cudaGetLastError();
    return PG_ERR_CUDA;
  }
  const int tiles = cdiv(p.M, SK_BM) * cdiv(p.N, BN);
  return launch_tc("gemm_tcgen05_splitk", kern, dim3(tiles * S), dim3(256), smem, S, true, st, ma, mw, p);
}

template <typename T, int EPI>
static int launch_sk_pick(const CUtensorMap& ma, const CUtensorMap& mw, const SkParams& p, int bn, int s, cudaStream_t st) {
  if (bn == 64) return s == 4 ? launch_sk<T, EPI, 64, 4>(ma, mw, p, st) : launch_sk<T, EPI, 64, 2>(ma, mw, p, st);
  return s == 4 ? launch_sk<T, EPI, 128, 4>(ma, mw, p, st) : launch_sk<T, EPI, 128, 2>(ma, mw, p, st);
}

}  // namespace tc

// Split factor for a problem (0 = do not split).  Chosen so tiles*S fills about two waves of 148 SMs and
// every CTA still has >= 8 K blocks to stream.
int gemm_tc_splitk_factor(int M, int N, int K, int epi, int* bn_out) {
  // PG_SPLITK: 0 never, 1 always when it applies, default (-1): only for a single row of tiles (M <= 128, the
  // batched-decode GEMMs): there W is read once anyway and the only problem is that 16-40 CTAs cannot pull
  // 6 TB/s.  With several m-tiles (260-token prefill) the GEMM is L2->SM bound and splitting measured 6 % slower.
  static const int mode = env_int("PG_SPLITK", -1);
  if (mode == 0 || epi == PG_EPI_GEGLU) return 0;
  const int kb = cdiv(K, tc::SK_BK);
  if (mode < 0 && M > tc::SK_BM) {
    // Exception: SigLIP at batch 1 (256 patch rows): out_proj and fc2 are 256 x 1152 outputs -- 36 CTAs of 64 columns,
    // fc2 with 68 serial K blocks each (28 us for a 10 MB matrix; ncu: 36 SMs active, ~640 cycles per K block).
    // Clusters must fit ONE wave: at most 4 clusters of 4 per GPC (36 clusters of 4 took two waves and 29 us).
    // Measured, SigLIP encode of one image: unsplit 2.009 ms; S=2 x 64 columns 1.803; S=4 x 128 columns 1.855; S=2 x
    // 128 columns 1.909; also splitting out_proj (20 K blocks) 1.935 -- that one sits on the launch floor already.
    static const int small_s = env_int("PG_SPLITK_SMALL_S", 2), small_bn = env_int("PG_SPLITK_SMALL_BN", 64);
    static const int small_kb = env_int("PG_SPLITK_SMALL_KB", 32);
    if (M > 2 * tc::SK_BM || N > 2048 || kb < small_kb || small_s < 2) return 0;
    const int tiles = cdiv(M, tc::SK_BM) * cdiv(N, small_bn);
    int s = small_s >= 4 ? 4 : 2;
    while (s >= 2 && (kb / s < 8 || tiles * s > (s == 4 ? 112 : 144))) s /= 2;
    if (s < 2) return 0;
    *bn_out = small_bn;
    return s;
  }
  const int bn = N <= 4096 ? 64 : 128;
  const int tiles = cdiv(M, tc::SK_BM) * cdiv(N, bn);
  if (tiles >= 120) return 0;
  int s = (tiles * 4 <= 320 && kb >= 32) ? 4 : ((kb >= 16) ? 2 : 0);
  *bn_out = bn;
  return s;
}

int gemm_tc_splitk(void* C, const void* A, const void* W, const void* bias, const void* R, int M, int N, int K, int lda,
                   int ldw, int ldc, int ldr, int res_mod, int epi, int out_f32, int dtype, int bn, int s,
                   cudaStream_t st) {
  const bool bf = dtype == PG_BF16;
  CUtensorMap ma, mw;
  PG_REQUIRE(tc::make_map_2d(&ma, A, M, K, lda, tc::SK_BM, bf) && tc::make_map_2d(&mw, W, N, K, ldw, bn, bf),
             "gemm_tc_splitk: cuTensorMapEncodeTiled failed");
  tc::SkParams p = {C, bias, R, M, N, K, ldc, ldr, res_mod, out_f32};
#define PG_SK(E) return bf ? tc::launch_sk_pick<bf16, E>(ma, mw, p, bn, s, st) : tc::launch_sk_pick<f16, E>(ma, mw, p, bn, s, st)
  switch (epi) {
    case PG_EPI_NONE: PG_SK(PG_EPI_NONE);
    case PG_EPI_BIAS: PG_SK(PG_EPI_BIAS);
    case PG_EPI_BIAS_GELU: PG_SK(PG_EPI_BIAS_GELU);
    case PG_EPI_BIAS_RES: PG_SK(PG_EPI_BIAS_RES);
    case PG_EPI_RES: PG_SK(PG_EPI_RES);
  }
#undef PG_SK
  set_error("gemm_tc_splitk: bad epilogue %d", epi);
  return PG_ERR_INVALID;
}

}  // namespace pg
