// Prompt-sized GEMM (128 < M <= 512 token rows: the 260-token prefill and the cache-off recompute of the reference,
// modeling_gemma.py:357-382 called with q_len = N + t):  out[M,N] = X[M,K] W[N,K]^T, swap-AB.
//
// With a few hundred rows the row-major kernel pads M to 128-row tiles (260 -> 384: a third of the MMA work is
// zeros) and every SM re-reads W once per m-tile; the per-SM ingest from L2 (~40-60 B/clk) is what bounds it.
// Here the WEIGHT tile is the 128-row A operand and ALL tokens are the N dimension of the MMA (NT = 272 or 512
// columns of one TMEM accumulator, issued as an N=256 MMA plus an N=NT-256 MMA per 16-wide K step), so a weight byte
// enters an SM exactly once and the padding is 272/260.  One stage = 16 KB of W + NT*128 B of X.
//   S == 1 : persistent over the N/128 weight tiles (gate/up, all-position lm_head).  GeGLU: a tile holds 64 gate rows
//            and the 64 matching up rows; the up half of the accumulator crosses to the gate threads through smem.
//   S  > 1 : a cluster of S CTAs splits K for one weight tile (q/k/v, o_proj, down_proj have 16-20 tiles); every CTA
//            parks its partial D^T in shared memory and then reduces + finishes ITS share of the token chunks
//            through distributed shared memory (no leader bottleneck).
// Accumulator is D^T (lane = output feature, column = token): bias is a per-thread scalar, stores of one token are
// contiguous across the warp.  Same rounding points as gemm_tc_kernel / gemm_simt.cu.
#include "tc_common.cuh"

namespace pg {
namespace tc {

constexpr int WD_BM = 128, WD_BK = 64;

struct WdParams {
  void* C;
  const void* bias;
  const void* R;
  int M, N, K, ldc, ldr, out_f32;
  int rotate;  // start each CTA's K loop at a different block (wrapping): every CTA reads the SAME token slices
};

__device__ __forceinline__ void wd_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float4 wd_ld_dsmem_v4(uint32_t local_addr, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(rank));
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(remote) : "memory");
  return v;
}
__device__ __forceinline__ void wd_epi_sync() { asm volatile("bar.sync 2, 128;" ::: "memory"); }

// NT: token columns of the accumulator (272 or 512).  S: K split (cluster size).
template <typename T, int EPI, int NT, int S>
__global__ void __launch_bounds__(256, 1)
gemm_tc_wide_kernel(const __grid_constant__ CUtensorMap map_x1, const __grid_constant__ CUtensorMap map_x2,
                    const __grid_constant__ CUtensorMap map_w, WdParams p) {
  constexpr bool DUAL = (EPI == PG_EPI_GEGLU);
  static_assert(!DUAL || S == 1, "GeGLU tiles are not K-split");
  constexpr int N1 = 256, N2 = NT - 256;                          // the two MMAs of a K step
  static_assert(N2 >= 16 && N2 <= 256 && N2 % 16 == 0, "unsupported token tile");
  constexpr int W_BYTES = WD_BM * WD_BK * 2, X1_BYTES = N1 * WD_BK * 2, X2_BYTES = N2 * WD_BK * 2;
  constexpr int X2_PAD = (X2_BYTES + 1023) / 1024 * 1024;
  constexpr int STAGE_BYTES = W_BYTES + X1_BYTES + X2_PAD;
  constexpr int NSTAGES = (200 * 1024) / STAGE_BYTES;
  static_assert(NSTAGES >= 2, "stage ring too small");
  constexpr int FEAT = DUAL ? 64 : WD_BM;                         // output features per weight tile
  constexpr uint32_t FMT = std::is_same<T, bf16>::value ? 1 : 0;
  constexpr uint32_t IDESC1 = umma_idesc(FMT, WD_BM, N1), IDESC2 = umma_idesc(FMT, WD_BM, N2);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = smem_base + NSTAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (NSTAGES + s); };
  const uint32_t tfull_bar = bars + 8u * (2 * NSTAGES), tempty_bar = tfull_bar + 8;
  const uint32_t tmem_slot = tfull_bar + 16;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  // after the main loop the ring is idle: S > 1 parks the fp32 partial there ([128][NT+4]); GeGLU (S == 1) uses a
  // small buffer BEHIND the barriers for the up half of a 16-token chunk ([64][20] floats)
  constexpr int NPASS = NT > 272 ? 2 : 1;                         // NT = 512: the partial is exchanged in two halves
  constexpr int PASS_CHUNKS = (NT / 16 + NPASS - 1) / NPASS;
  constexpr int PROW = PASS_CHUNKS * 16 + 4;
  static_assert(WD_BM * PROW * 4 <= NSTAGES * STAGE_BYTES, "partial tile does not fit the idle stage ring");
  float* part = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)));
  float* ubuf = reinterpret_cast<float*>(smem_raw + (bars - smem_u32(smem_raw)) + 8 * (2 * NSTAGES + 2) + 16);

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t crank = 0;
  if (S > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const int n_tiles = (p.N + FEAT - 1) / FEAT;
  const int k_blocks = (p.K + WD_BK - 1) / WD_BK, per = (k_blocks + S - 1) / S;
  const int kb0 = (int)crank * per, kb1 = min(k_blocks, kb0 + per);
  const int tile_first = (S > 1) ? (int)(blockIdx.x / S) : (int)blockIdx.x;
  const int tile_stride = (S > 1) ? n_tiles : (int)gridDim.x;      // S > 1: exactly one tile per cluster
  const int nk = kb1 - kb0;                                        // K blocks of this CTA, visited from kb0 + rot, wrapping
  const int rot = (p.rotate && nk > 1) ? (int)((blockIdx.x * 7u) % (unsigned)nk) : 0;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NSTAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer: weight tile + every token's K slice =====================
    int stage = 0;
    uint32_t phase = 0;
    pdl_wait();
    for (int tile = tile_first; tile < n_tiles; tile += tile_stride) {
      for (int i = 0; i < nk; ++i) {
        const int kb = kb0 + (i + rot < nk ? i + rot : i + rot - nk);
        mbar_wait(empty_bar(stage), phase ^ 1);
        const uint32_t sa = smem_base + stage * STAGE_BYTES;
        mbar_expect_tx(full_bar(stage), W_BYTES + X1_BYTES + X2_BYTES);
        if (DUAL) {  // 64 gate rows, then the 64 matching up rows (map_w has a 64-row box)
          tma_load_2d(sa, &map_w, full_bar(stage), kb * WD_BK, tile * 64);
          tma_load_2d(sa + W_BYTES / 2, &map_w, full_bar(stage), kb * WD_BK, p.N + tile * 64);
        } else {
          tma_load_2d(sa, &map_w, full_bar(stage), kb * WD_BK, tile * WD_BM);
        }
        tma_load_2d(sa + W_BYTES, &map_x1, full_bar(stage), kb * WD_BK, 0);
        tma_load_2d(sa + W_BYTES + X1_BYTES, &map_x2, full_bar(stage), kb * WD_BK, N1);
        if (++stage == NSTAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer: D^T[128 features, NT tokens] += W_tile X^T =====================
    int stage = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int tile = tile_first; tile < n_tiles; tile += tile_stride) {
      mbar_wait(tempty_bar, acc_phase ^ 1);   // the epilogue has drained the (single) accumulator
      tc_fence_after();
      for (int i = 0; i < nk; ++i) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * STAGE_BYTES;
#pragma unroll
        for (int k = 0; k < WD_BK / 16; ++k) {
          const uint64_t wd = umma_desc(sa + k * 32);
          const uint32_t accumulate = (i > 0 || k > 0) ? 1u : 0u;
          umma(tmem_base, wd, umma_desc(sa + W_BYTES + k * 32), IDESC1, accumulate);
          umma(tmem_base + N1, wd, umma_desc(sa + W_BYTES + X1_BYTES + k * 32), IDESC2, accumulate);
        }
        umma_commit(empty_bar(stage));
        if (++stage == NSTAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(tfull_bar);
      acc_phase ^= 1;
    }
  }
  __syncwarp();

  // ===================== epilogue: thread <-> output feature (TMEM lane), columns <-> tokens =====================
  const int q = warp & 3, r = q * 32 + lane;
  const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
  T* Ct = reinterpret_cast<T*>(p.C);
  float* Cf = reinterpret_cast<float*>(p.C);
  const T* bias = reinterpret_cast<const T*>(p.bias);
  const T* R = reinterpret_cast<const T*>(p.R);
  // 16 tokens of one output feature n (v: accumulators; u: the up half for GeGLU)
  auto finish = [&](int n, const float* v, const float* u, int m0) {
    if (n >= p.N) return;
    float bn = 0.f;
    if (EPI == PG_EPI_BIAS || EPI == PG_EPI_BIAS_GELU || EPI == PG_EPI_BIAS_RES) bn = to_f<T>(bias[n]);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int m = m0 + j;
      if (m >= p.M) break;
      float x = rnd<T>(v[j] + bn);
      if (EPI == PG_EPI_BIAS_GELU) x = rnd<T>(gelu_tanh_fast(x));
      if (EPI == PG_EPI_GEGLU) x = rnd<T>(rnd<T>(gelu_tanh_fast(x)) * rnd<T>(u[j]));
      if (EPI == PG_EPI_BIAS_RES || EPI == PG_EPI_RES) x = rnd<T>(x + to_f<T>(R[(size_t)m * p.ldr + n]));
      if (p.out_f32) Cf[(size_t)m * p.ldc + n] = x;
      else Ct[(size_t)m * p.ldc + n] = from_f<T>(x);
    }
  };
  const int m_chunks = (p.M + 15) / 16;   // 16-token chunks that hold real rows

  if constexpr (S == 1) {
    if (warp >= 4) {
      uint32_t acc_phase = 0;
      pdl_wait();
      for (int tile = tile_first; tile < n_tiles; tile += tile_stride) {
        mbar_wait(tfull_bar, acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < m_chunks; ++c) {
          float v[16];
          tmem_ld16(t_lane + c * 16, v);
          tmem_ld_wait();
          if (DUAL) {
            // lanes 64..127 hold the up projections of features tile*64 + (r - 64): hand them to the gate threads
            if (q >= 2) {
#pragma unroll
              for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<float4*>(&ubuf[(r - 64) * 20 + j]) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
            wd_epi_sync();
            if (q < 2) {
              float u[16];
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                const float4 t4 = *reinterpret_cast<const float4*>(&ubuf[r * 20 + j]);
                u[j] = t4.x; u[j + 1] = t4.y; u[j + 2] = t4.z; u[j + 3] = t4.w;
              }
              finish(tile * 64 + r, v, u, c * 16);
            }
            wd_epi_sync();   // ubuf is rewritten by the next chunk
          } else {
            finish(tile * WD_BM + r, v, v, c * 16);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar);
        acc_phase ^= 1;
      }
    }
  } else {
    const int tile = tile_first;
    const bool have_k = kb1 > kb0;
    if (warp >= 4) {
      pdl_wait();
      mbar_wait(tfull_bar, 0);   // this CTA's MMAs are complete: accumulator final, stage ring idle
      tc_fence_after();
    }
#pragma unroll 1
    for (int pass = 0; pass < NPASS; ++pass) {
      const int c_lo = pass * PASS_CHUNKS, c_hi = min(m_chunks, c_lo + PASS_CHUNKS);
      if (warp >= 4) {
#pragma unroll 1
        for (int c = c_lo; c < c_hi; ++c) {
          float v[16];
          tmem_ld16(t_lane + c * 16, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(&part[r * PROW + (c - c_lo) * 16 + j]) =
                have_k ? make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      tc_fence_before();
      wd_cluster_sync();
      if (warp >= 4) {
        // this CTA finishes the token chunks c == crank (mod S): sum the S partials in rank order (deterministic)
#pragma unroll 1
        for (int c = c_lo + (int)crank; c < c_hi; c += S) {
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0.f;
#pragma unroll 1
          for (int peer = 0; peer < S; ++peer) {
            float4 pv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
              pv[j] = wd_ld_dsmem_v4(smem_base + (uint32_t)((r * PROW + (c - c_lo) * 16 + 4 * j) * 4), peer);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              v[4 * j] += pv[j].x; v[4 * j + 1] += pv[j].y; v[4 * j + 2] += pv[j].z; v[4 * j + 3] += pv[j].w;
            }
          }
          finish(tile * WD_BM + r, v, v, c * 16);
        }
      }
      wd_cluster_sync();   // nobody overwrites or leaves while a peer may still read its partial
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

template <typename T, int EPI, int NT, int S>
static int launch_wd(const CUtensorMap& mx1, const CUtensorMap& mx2, const CUtensorMap& mw, const WdParams& p, cudaStream_t st) {
  constexpr bool DUAL = (EPI == PG_EPI_GEGLU);
  constexpr int X2_PAD = ((NT - 256) * WD_BK * 2 + 1023) / 1024 * 1024;
  constexpr int STAGE_BYTES = WD_BM * WD_BK * 2 + 256 * WD_BK * 2 + X2_PAD;
  constexpr int NSTAGES = (200 * 1024) / STAGE_BYTES;
  const size_t smem = 1024 + (size_t)NSTAGES * STAGE_BYTES + 8 * (2 * NSTAGES + 2) + 16 + (DUAL ? 64 * 20 * 4 : 0) + 16;
  auto kern = gemm_tc_wide_kernel<T, EPI, NT, S>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    set_error("gemm_tc_wide: cannot reserve %zu B of shared memory", smem);
    cudaGetLastError();
    return PG_ERR_CUDA;
  }
  const int n_tiles = cdiv(p.N, DUAL ? 64 : WD_BM);
  const int grid = S > 1 ? n_tiles * S : (n_tiles < 148 ? n_tiles : 148);
  return launch_tc("gemm_tcgen05_wide", kern, dim3(grid), dim3(256), smem, S, true, st, mx1, mx2, mw, p);
}

template <typename T, int EPI, int NT>
static int launch_wd_s(const CUtensorMap& mx1, const CUtensorMap& mx2, const CUtensorMap& mw, const WdParams& p, int s,
                       cudaStream_t st) {
  if constexpr (EPI != PG_EPI_GEGLU) {
    if (s == 8) return launch_wd<T, EPI, NT, 8>(mx1, mx2, mw, p, st);
    if (s == 4) return launch_wd<T, EPI, NT, 4>(mx1, mx2, mw, p, st);
    if (s == 2) return launch_wd<T, EPI, NT, 2>(mx1, mx2, mw, p, st);
  }
  return launch_wd<T, EPI, NT, 1>(mx1, mx2, mw, p, st);
}

}  // namespace tc

// Prompt-sized problems: 128 < M <= 512 rows.  OFF by default (PG_WIDE=1 or pg_gemm impl=3 select it): measured on the
// 260-token prefill it is slower than the row-major kernels (gate/up 81 vs 55 us, down 104 vs 85, qkv 26 vs 14, o_proj
// 80 vs 10 us per layer).  Every CTA re-reads all tokens, so the L2->SM ingest (~30 B/clk/SM with all 148 SMs pulling)
// stays the bound, 256 tiles over 148 CTAs quantise to two rounds, the single 272-column accumulator cannot overlap
// the epilogue with the next tile, and 8-CTA clusters of 205 KB CTAs do not all fit at once.  The version that would
// win shares the token slices across a CTA pair (cta_group::2 with the tokens as the split B operand).
bool gemm_tc_wide_supported(int M, int N, int K, int epi) {
  return M > 128 && M <= 512 && N >= 512 && K >= 256 && epi >= PG_EPI_NONE && epi <= PG_EPI_GEGLU;
}
bool gemm_tc_wide_wanted(int M, int N, int K, int epi) {
  static const int enabled = env_int("PG_WIDE", 0);
  return enabled && gemm_tc_wide_supported(M, N, K, epi);
}

int gemm_tc_wide(void* C, const void* A, const void* W, const void* bias, const void* R, int M, int N, int K, int lda,
                 int ldw, int ldc, int ldr, int epi, int out_f32, int dtype, cudaStream_t st) {
  const bool bf = dtype == PG_BF16;
  const int nt = M <= 272 ? 272 : 512;
  const bool dual = epi == PG_EPI_GEGLU;
  const int n_tiles = cdiv(N, dual ? 64 : tc::WD_BM), kb = cdiv(K, tc::WD_BK);
  // split K across a cluster until ~148 CTAs stream (each keeps >= 4 K blocks)
  static const int s_env = env_int("PG_WIDE_S", 0);
  int s = 1;
  if (!dual) {
    while (s < 4 && n_tiles * s * 2 <= 148 && kb / (s * 2) >= 4) s *= 2;  // 8-CTA clusters of 205 KB CTAs do not all fit
    if (s_env) s = s_env;
  }
  CUtensorMap mx1, mx2, mw;
  const int w_rows = dual ? 2 * N : N;
  PG_REQUIRE(tc::make_map_2d(&mx1, A, M, K, lda, 256, bf) && tc::make_map_2d(&mx2, A, M, K, lda, nt - 256, bf) &&
                 tc::make_map_2d(&mw, W, w_rows, K, ldw, dual ? 64 : tc::WD_BM, bf),
             "gemm_tc_wide: cuTensorMapEncodeTiled failed");
  static const int rot_env = env_int("PG_WIDE_ROT", 1);
  tc::WdParams p = {C, bias, R, M, N, K, ldc, ldr, out_f32, rot_env};
#define PG_WD(E)                                                                                                  \
  if (nt == 272) return bf ? tc::launch_wd_s<bf16, E, 272>(mx1, mx2, mw, p, s, st) : tc::launch_wd_s<f16, E, 272>(mx1, mx2, mw, p, s, st); \
  return bf ? tc::launch_wd_s<bf16, E, 512>(mx1, mx2, mw, p, s, st) : tc::launch_wd_s<f16, E, 512>(mx1, mx2, mw, p, s, st)
  switch (epi) {
    case PG_EPI_NONE: PG_WD(PG_EPI_NONE);
    case PG_EPI_BIAS: PG_WD(PG_EPI_BIAS);
    case PG_EPI_BIAS_GELU: PG_WD(PG_EPI_BIAS_GELU);
    case PG_EPI_BIAS_RES: PG_WD(PG_EPI_BIAS_RES);
    case PG_EPI_RES: PG_WD(PG_EPI_RES);
    case PG_EPI_GEGLU: PG_WD(PG_EPI_GEGLU);
  }
#undef PG_WD
  set_error("gemm_tc_wide: bad epilogue %d", epi);
  return PG_ERR_INVALID;
}

}  // namespace pg
