// Skinny GEMM for batched decode (4 <= M <= 128 token rows, BASELINE configs[3]): out[M,N] = X[M,K] W[N,K]^T.
// With so few rows the problem is a weight stream, so the operands are swapped: the WEIGHT tile is the 128-row
// A operand of tcgen05.mma and the tokens are the N dimension (N_mma = M rounded up to 16).  A pipeline stage is
// then 16 KB of weights + a few KB of activations, so 8-10 stages (~150 KB of weights) are in flight per SM —
// what a 6 TB/s stream needs — instead of the 16 KB-activation/8 KB-weight stages of the row-major kernel.
// The accumulator is D^T: TMEM lane = output feature, column = token, so bias is a per-thread scalar and the
// stores are coalesced across the warp (consecutive features of one token).
//   S == 1 : persistent over the N/128 weight tiles (gate/up with the GeGLU pairing, lm_head), double-buffered
//            accumulators so the epilogue overlaps the next tile's stream;
//   S  > 1 : a cluster of S CTAs splits K for one weight tile (q/k/v, o_proj, down_proj have 16-20 tiles only);
//            partial D^T tiles meet in the leader through distributed shared memory.
#include "tc_common.cuh"

namespace pg {
namespace tc {

constexpr int SN_BM = 128, SN_BK = 64;

struct SnParams {
  void* C;
  const void* bias;
  const void* R;
  int M, N, K, ldc, ldr, out_f32;
};

__device__ __forceinline__ void sn_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float4 sn_ld_dsmem_v4(uint32_t local_addr, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(rank));
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(remote) : "memory");
  return v;
}

// NT: token columns of the accumulator (multiple of 16, <= 256).  S: K split (cluster size).
template <typename T, int EPI, int NT, int S>
__global__ void __launch_bounds__(256, 1)
gemm_tc_skinny_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, SnParams p) {
  constexpr bool DUAL = (EPI == PG_EPI_GEGLU);
  constexpr int NW = DUAL ? 2 : 1;                               // weight tiles per stage
  constexpr int W_BYTES = SN_BM * SN_BK * 2, X_BYTES = NT * SN_BK * 2;
  constexpr int X_PAD = (X_BYTES + 1023) / 1024 * 1024;          // keep every tile 1024-byte aligned
  constexpr int STAGE_BYTES = NW * W_BYTES + X_PAD;
  constexpr int NSTAGES = (200 * 1024) / STAGE_BYTES > 10 ? 10 : (200 * 1024) / STAGE_BYTES;
  constexpr int ACC_COLS = NW * NT;                              // TMEM columns per accumulator stage
  constexpr int TMEM_COLS = 2 * ACC_COLS <= 32 ? 32 : (2 * ACC_COLS <= 64 ? 64 : (2 * ACC_COLS <= 128 ? 128 : (2 * ACC_COLS <= 256 ? 256 : 512)));
  static_assert(2 * ACC_COLS <= 512, "accumulators do not fit TMEM");
  constexpr uint32_t IDESC = umma_idesc(std::is_same<T, bf16>::value ? 1 : 0, SN_BM, NT);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = smem_base + NSTAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (NSTAGES + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * NSTAGES + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * NSTAGES + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * NSTAGES + 4);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  // S > 1: fp32 partial D^T, one padded row per feature ([128][NT+4]: 16-byte accesses, conflict-free), reusing the ring
  constexpr int PROW = NT + 4;
  float* part = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)));

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t crank = 0;
  if (S > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const int n_tiles = (p.N + SN_BM - 1) / SN_BM;
  const int k_blocks = (p.K + SN_BK - 1) / SN_BK, per = (k_blocks + S - 1) / S;
  const int kb0 = (int)crank * per, kb1 = min(k_blocks, kb0 + per);
  const int tile_first = (S > 1) ? (int)(blockIdx.x / S) : (int)blockIdx.x;
  const int tile_stride = (S > 1) ? n_tiles : (int)gridDim.x;      // S > 1: exactly one tile per cluster

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NSTAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
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
    // ===================== TMA producer: weight tile(s) + the token block =====================
    // flat item index i -> (tile, k block).  The first NSTAGES items' WEIGHT tiles are requested before
    // griddepcontrol.wait: weights never depend on the preceding kernel, so ~150 KB per SM is already in flight
    // when the predecessor's activations become visible; the token blocks follow the wait.
    const int nk = kb1 - kb0;
    const int my_tiles = tile_first < n_tiles ? (n_tiles - tile_first + tile_stride - 1) / tile_stride : 0;
    const int items = my_tiles * nk;
    auto item_tile = [&](int i) { return tile_first + (i / nk) * tile_stride; };
    auto item_kb = [&](int i) { return kb0 + i % nk; };
    const int pre = items < NSTAGES ? items : NSTAGES;
    for (int i = 0; i < pre; ++i) {                               // fresh barriers: every stage is free
      const uint32_t sa = smem_base + i * STAGE_BYTES;
      mbar_expect_tx(full_bar(i), NW * W_BYTES + X_BYTES);
      tma_load_2d(sa, &map_w, full_bar(i), item_kb(i) * SN_BK, item_tile(i) * SN_BM);
      if (DUAL) tma_load_2d(sa + W_BYTES, &map_w, full_bar(i), item_kb(i) * SN_BK, p.N + item_tile(i) * SN_BM);
    }
    pdl_wait();
    for (int i = 0; i < pre; ++i)
      tma_load_2d(smem_base + i * STAGE_BYTES + NW * W_BYTES, &map_x, full_bar(i), item_kb(i) * SN_BK, 0);
    int stage = pre == NSTAGES ? 0 : pre;
    uint32_t phase = pre == NSTAGES ? 1 : 0;
    for (int i = pre; i < items; ++i) {
      mbar_wait(empty_bar(stage), phase ^ 1);
      const uint32_t sa = smem_base + stage * STAGE_BYTES;
      mbar_expect_tx(full_bar(stage), NW * W_BYTES + X_BYTES);
      tma_load_2d(sa, &map_w, full_bar(stage), item_kb(i) * SN_BK, item_tile(i) * SN_BM);
      if (DUAL) tma_load_2d(sa + W_BYTES, &map_w, full_bar(stage), item_kb(i) * SN_BK, p.N + item_tile(i) * SN_BM);
      tma_load_2d(sa + NW * W_BYTES, &map_x, full_bar(stage), item_kb(i) * SN_BK, 0);
      if (++stage == NSTAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer: D^T[128 features, NT tokens] += W_tile X_tile^T =====================
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int tile = tile_first; tile < n_tiles; tile += tile_stride) {
      mbar_wait(tempty_bar(acc), acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * STAGE_BYTES;
#pragma unroll
        for (int k = 0; k < SN_BK / 16; ++k) {
          const uint64_t xd = umma_desc(sa + NW * W_BYTES + k * 32);
          const uint32_t accumulate = (kb > kb0 || k > 0) ? 1u : 0u;
          umma(d_tmem, umma_desc(sa + k * 32), xd, IDESC, accumulate);
          if (DUAL) umma(d_tmem + NT, umma_desc(sa + W_BYTES + k * 32), xd, IDESC, accumulate);
        }
        umma_commit(empty_bar(stage));
        if (++stage == NSTAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(tfull_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  __syncwarp();

  // ===================== epilogue: thread <-> output feature, columns <-> tokens =====================
  const int q = warp & 3, r = q * 32 + lane;
  const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
  T* Ct = reinterpret_cast<T*>(p.C);
  float* Cf = reinterpret_cast<float*>(p.C);
  const T* bias = reinterpret_cast<const T*>(p.bias);
  const T* R = reinterpret_cast<const T*>(p.R);
  if (warp >= 4) pdl_wait();   // residual reads / output writes only after the predecessor has finished
  auto finish = [&](int tile, float (&v)[16], float (&u)[DUAL ? 16 : 1], int m0) {
    const int n = tile * SN_BM + r;
    if (n >= p.N) return;
    float bn = 0.f;
    if (EPI == PG_EPI_BIAS || EPI == PG_EPI_BIAS_GELU || EPI == PG_EPI_BIAS_RES) bn = to_f<T>(bias[n]);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int m = m0 + j;
      if (m >= p.M) break;
      float x = rnd<T>(v[j] + bn);
      if (EPI == PG_EPI_BIAS_GELU) x = rnd<T>(gelu_tanh_fast(x));
      if (EPI == PG_EPI_GEGLU) x = rnd<T>(rnd<T>(gelu_tanh_fast(x)) * rnd<T>(u[DUAL ? j : 0]));
      if (EPI == PG_EPI_BIAS_RES || EPI == PG_EPI_RES) x = rnd<T>(x + to_f<T>(R[(size_t)m * p.ldr + n]));
      if (p.out_f32) Cf[(size_t)m * p.ldc + n] = x;
      else Ct[(size_t)m * p.ldc + n] = from_f<T>(x);
    }
  };

  if constexpr (S == 1) {
    if (warp >= 4) {
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = tile_first; tile < n_tiles; tile += tile_stride) {
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < NT / 16; ++c) {
          float v[16], u[DUAL ? 16 : 1];
          tmem_ld16(t_lane + acc * ACC_COLS + c * 16, v);
          if (DUAL) tmem_ld16(t_lane + acc * ACC_COLS + NT + c * 16, u);
          tmem_ld_wait();
          finish(tile, v, u, c * 16);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    const int tile = tile_first;
    const bool have_k = kb1 > kb0;
    if (warp >= 4) {
      mbar_wait(tfull_bar(0), 0);   // this CTA's MMAs are complete: accumulator final, stage ring idle
      tc_fence_after();
      if (crank != 0) {
#pragma unroll 1
        for (int c = 0; c < NT / 16; ++c) {
          float v[16];
          tmem_ld16(t_lane + c * 16, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(&part[r * PROW + c * 16 + j]) =
                have_k ? make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
    tc_fence_before();
    sn_cluster_sync();
    if (crank == 0 && warp >= 4) {
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < NT / 16; ++c) {
        float v[16], u[DUAL ? 16 : 1];
        tmem_ld16(t_lane + c * 16, v);
        tmem_ld_wait();
        float4 pv[S > 1 ? S - 1 : 1][4];
#pragma unroll
        for (int peer = 1; peer < S; ++peer)   // all remote loads in flight before the first add
#pragma unroll
          for (int j = 0; j < 4; ++j)
            pv[peer - 1][j] = sn_ld_dsmem_v4(smem_base + (uint32_t)((r * PROW + c * 16 + 4 * j) * 4), peer);
#pragma unroll
        for (int peer = 1; peer < S; ++peer)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            v[4 * j] += pv[peer - 1][j].x; v[4 * j + 1] += pv[peer - 1][j].y;
            v[4 * j + 2] += pv[peer - 1][j].z; v[4 * j + 3] += pv[peer - 1][j].w;
          }
        finish(tile, v, u, c * 16);
      }
    }
    tc_fence_before();
    sn_cluster_sync();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

template <typename T, int EPI, int NT, int S>
static int launch_sn(const CUtensorMap& mx, const CUtensorMap& mw, const SnParams& p, cudaStream_t st) {
  constexpr bool DUAL = (EPI == PG_EPI_GEGLU);
  constexpr int X_PAD = (NT * SN_BK * 2 + 1023) / 1024 * 1024;
  constexpr int STAGE_BYTES = (DUAL ? 2 : 1) * SN_BM * SN_BK * 2 + X_PAD;
  constexpr int NSTAGES = (200 * 1024) / STAGE_BYTES > 10 ? 10 : (200 * 1024) / STAGE_BYTES;
  const size_t smem = 1024 + (size_t)NSTAGES * STAGE_BYTES + 8 * (2 * NSTAGES + 4) + 16;
  auto kern = gemm_tc_skinny_kernel<T, EPI, NT, S>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    set_error("gemm_tc_skinny: cannot reserve %zu B of shared memory", smem);
    cudaGetLastError();
    return PG_ERR_CUDA;
  }
  const int n_tiles = cdiv(p.N, SN_BM);
  const int grid = S > 1 ? n_tiles * S : (n_tiles < 148 ? n_tiles : 148);
  return launch_tc("gemm_tcgen05_skinny", kern, dim3(grid), dim3(256), smem, S, true, st, mx, mw, p);
}

template <typename T, int EPI, int NT>
static int launch_sn_s(const CUtensorMap& mx, const CUtensorMap& mw, const SnParams& p, int s, cudaStream_t st) {
  if (EPI != PG_EPI_GEGLU) {
    if (s == 8) return launch_sn<T, (EPI == PG_EPI_GEGLU ? PG_EPI_NONE : EPI), NT, 8>(mx, mw, p, st);
    if (s == 4) return launch_sn<T, (EPI == PG_EPI_GEGLU ? PG_EPI_NONE : EPI), NT, 4>(mx, mw, p, st);
    if (s == 2) return launch_sn<T, (EPI == PG_EPI_GEGLU ? PG_EPI_NONE : EPI), NT, 2>(mx, mw, p, st);
  }
  return launch_sn<T, EPI, NT, 1>(mx, mw, p, st);
}

template <typename T, int EPI>
static int launch_sn_nt(const CUtensorMap& mx, const CUtensorMap& mw, const SnParams& p, int nt, int s, cudaStream_t st) {
  if (nt <= 32) return launch_sn_s<T, EPI, 32>(mx, mw, p, s, st);
  if (nt <= 64) return launch_sn_s<T, EPI, 64>(mx, mw, p, s, st);
  return launch_sn_s<T, EPI, 128>(mx, mw, p, s, st);
}

}  // namespace tc

bool gemm_tc_skinny_wanted(int M, int N, int K, int epi) {
  // Measured on the batch-32 decode step (tools/kernel_sweep.py, SWEEP_B=32): the swap-AB stream wins where a CTA
  // owns whole weight tiles (gate/up 26.6 vs 29.0 us, lm_head 181 vs 211 us) and for q/k/v (9.9 vs 18.3 us); with a
  // residual epilogue and a K split (o_proj, down_proj: 16 tiles) the row-major split-K kernel is faster.
  static const int mode = env_int("PG_SKINNY", 1);   // 0 off, 1 where it wins, 2 everywhere it applies
  if (!mode || M < 4 || M > 128 || N < 512 || K < 256 || epi < PG_EPI_NONE || epi > PG_EPI_GEGLU) return false;
  if (mode == 2) return true;
  const int n_tiles = cdiv(N, 128);
  // up to 8 rows the swap-AB stream also wins for the residual projections (batch 8 step: 1.92 vs 2.00 ms)
  return epi == PG_EPI_GEGLU || epi == PG_EPI_NONE || n_tiles >= 100 || M <= 8;
}

int gemm_tc_skinny(void* C, const void* A, const void* W, const void* bias, const void* R, int M, int N, int K, int lda,
                   int ldw, int ldc, int ldr, int epi, int out_f32, int dtype, cudaStream_t st) {
  const bool bf = dtype == PG_BF16;
  const int nt = M <= 32 ? 32 : (M <= 64 ? 64 : 128);
  const int n_tiles = cdiv(N, tc::SN_BM), kb = cdiv(K, tc::SN_BK);
  // split K across a cluster until ~148 CTAs stream (each keeps >= 8 K blocks)
  int s = 1;
  if (epi != PG_EPI_GEGLU)
    while (s < 4 && n_tiles * s * 2 <= 160 && kb / (s * 2) >= 8) s *= 2;
  CUtensorMap mx, mw;
  const int w_rows = (epi == PG_EPI_GEGLU) ? 2 * N : N;
  PG_REQUIRE(tc::make_map_2d(&mx, A, M, K, lda, nt, bf) && tc::make_map_2d(&mw, W, w_rows, K, ldw, tc::SN_BM, bf),
             "gemm_tc_skinny: cuTensorMapEncodeTiled failed");
  tc::SnParams p = {C, bias, R, M, N, K, ldc, ldr, out_f32};
#define PG_SN(E) return bf ? tc::launch_sn_nt<bf16, E>(mx, mw, p, nt, s, st) : tc::launch_sn_nt<f16, E>(mx, mw, p, nt, s, st)
  switch (epi) {
    case PG_EPI_NONE: PG_SN(PG_EPI_NONE);
    case PG_EPI_BIAS: PG_SN(PG_EPI_BIAS);
    case PG_EPI_BIAS_GELU: PG_SN(PG_EPI_BIAS_GELU);
    case PG_EPI_BIAS_RES: PG_SN(PG_EPI_BIAS_RES);
    case PG_EPI_RES: PG_SN(PG_EPI_RES);
    case PG_EPI_GEGLU: PG_SN(PG_EPI_GEGLU);
  }
#undef PG_SN
  set_error("gemm_tc_skinny: bad epilogue %d", epi);
  return PG_ERR_INVALID;
}

}  // namespace pg
