// C[M,N] = A[M,K] W[N,K]^T on the 5th-generation tensor cores: TMA (cp.async.bulk.tensor,
// SWIZZLE_128B) stages A and W tiles in shared memory, one elected thread issues tcgen05.mma
// (cta_group::1, kind::f16, M=128 x N=128 x K=16) with fp32 accumulators in TMEM, four epilogue warps
// read them back with tcgen05.ld and apply the fused epilogue (bias / GELU / residual / GeGLU /
// fp32 out) with the same rounding points as the SIMT reference kernel (gemm_simt.cu).
//
// Persistent, warp-specialised, one CTA per SM:
//   warp 0  TMA producer      smem ring of NSTAGES x {A 128x64, W 128x64 [, W_up 128x64]}
//   warp 1  MMA issuer        4 x tcgen05.mma per stage, tcgen05.commit frees the stage
//   warp 2  TMEM allocator    512 columns = 2 accumulator stages x (1 or 2) x 128 columns
//   warps 4-7 epilogue        TMEM -> registers -> global, overlapped with the next tile's MMAs
// Used by the prefill / cache-off recompute / SigLIP / projector / lm_head(all positions) GEMMs of
// the reference (every nn.Linear of SURVEY.md §2.3 with more than a handful of rows).
#include "tc_common.cuh"

namespace pg {
namespace tc {

constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr int THREADS = 256;
constexpr int TMEM_COLS = 512;
constexpr int TILE_BYTES = BM * BK * 2;  // 16 KB (A tile; a W tile is BN*BK*2)

struct Params {
  void* C;
  const void* bias;
  const void* R;
  int M, N, K, ldc, ldr, res_mod, out_f32;
};

// CL = 2: CTA pairs (a 2-CTA cluster) work on vertically adjacent tiles (same n_blk, m_blk = 2*mp + rank): each CTA
// loads its own A tile and HALF of the shared W tile, multicast into both CTAs' shared memory, so the L2->SM
// traffic per FLOP drops by a third (the 128x256 single-CTA tile needs ~94 B/clk/SM, more than the fabric gives).
template <typename T, int EPI, int BN, int CL>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, Params p) {
  constexpr bool DUAL = (EPI == PG_EPI_GEGLU);
  constexpr int NB_TILES = DUAL ? 2 : 1;              // W tiles per stage
  constexpr int NSTAGES = (DUAL || BN > 128) ? 4 : 6;
  constexpr int W_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = TILE_BYTES + NB_TILES * W_BYTES;
  constexpr int ACC_COLS = NB_TILES * BN;             // TMEM columns per accumulator stage
  constexpr uint32_t IDESC = umma_idesc(std::is_same<T, bf16>::value ? 1 : 0, BM, BN);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-B alignment
  const uint32_t bars = smem_base + NSTAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (NSTAGES + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * NSTAGES + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * NSTAGES + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * NSTAGES + 4);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  pdl_launch_dependents();  // the next kernel of the chain may set itself up while this one runs
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles_real = (p.M + BM - 1) / BM, n_tiles = (p.N + BN - 1) / BN;
  // CL == 2: schedule over (n_blk, m-pair); the CTA's own tile is m_blk = 2*mp + rank (a ghost tile past M loads
  // zeros and stores nothing).  CL == 1: plain tiles.
  uint32_t crank = 0;
  if (CL == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const int m_tiles = (CL == 2) ? (m_tiles_real + 1) / 2 : m_tiles_real;   // schedule units along M
  const int total_tiles = m_tiles * n_tiles;
  const int sched_first = (CL == 2) ? (int)(blockIdx.x / 2) : (int)blockIdx.x;
  const int sched_stride = (CL == 2) ? (int)(gridDim.x / 2) : (int)gridDim.x;
  const int k_blocks = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NSTAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), CL); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (CL == 2) {  // the peer's barriers must exist before anything is multicast into / arrives on them
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      pdl_wait();  // A (and everything else) comes from the preceding kernels
      for (int tile = sched_first; tile < total_tiles; tile += sched_stride) {
        const int n_blk = tile / m_tiles;                          // neighbours share the W tile (L2 reuse)
        const int m_blk = (CL == 2) ? 2 * (tile % m_tiles) + (int)crank : tile % m_tiles;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);                  // CL == 2: both CTAs have consumed this stage
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          mbar_expect_tx(full_bar(stage), STAGE_BYTES);
          tma_load_2d(sa, &map_a, full_bar(stage), kb * BK, m_blk * BM);
          if (CL == 2) {  // this CTA's half of the W tile, delivered to both CTAs of the pair
            tma_load_2d_mcast(sa + TILE_BYTES + crank * (W_BYTES / 2), &map_w, full_bar(stage), kb * BK,
                              n_blk * BN + (int)crank * (BN / 2), (uint16_t)0x3);
          } else
          tma_load_2d(sa + TILE_BYTES, &map_w, full_bar(stage), kb * BK, n_blk * BN);
          if (DUAL) tma_load_2d(sa + TILE_BYTES + W_BYTES, &map_w, full_bar(stage), kb * BK, p.N + n_blk * BN);
          if (++stage == NSTAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = sched_first; tile < total_tiles; tile += sched_stride) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);  // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t ad = umma_desc(sa + k * UMMA_K * 2);
            const uint32_t accumulate = (kb > 0 || k > 0) ? 1u : 0u;
            umma(d_tmem, ad, umma_desc(sa + TILE_BYTES + k * UMMA_K * 2), IDESC, accumulate);
            if (DUAL) umma(d_tmem + BN, ad, umma_desc(sa + TILE_BYTES + W_BYTES + k * UMMA_K * 2), IDESC, accumulate);
          }
          if (CL == 2) umma_commit_mcast(empty_bar(stage), (uint16_t)0x3);  // frees the stage in BOTH CTAs
          else umma_commit(empty_bar(stage));  // smem stage reusable once these MMAs have read it
          if (++stage == NSTAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar(acc));  // accumulator complete
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM -> registers -> global =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    T* Ct = reinterpret_cast<T*>(p.C);
    float* Cf = reinterpret_cast<float*>(p.C);
    const T* bias = reinterpret_cast<const T*>(p.bias);
    const T* R = reinterpret_cast<const T*>(p.R);
    pdl_wait();  // residual reads and output writes only after the predecessor has finished
    for (int tile = sched_first; tile < total_tiles; tile += sched_stride) {
      const int n_blk = tile / m_tiles;
      const int m_blk = (CL == 2) ? 2 * (tile % m_tiles) + (int)crank : tile % m_tiles;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const int m = m_blk * BM + q * 32 + lane;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * ACC_COLS;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        float v[32], u[DUAL ? 32 : 1];
        tmem_ld32(t_row + c * 32, v);
        if (DUAL) tmem_ld32(t_row + BN + c * 32, u);
        tmem_ld_wait();
        const int n0 = n_blk * BN + c * 32;
        if (m < p.M && n0 < p.N) {
          const int rm = p.res_mod > 0 ? (m % p.res_mod) : m;
#pragma unroll
          for (int j0 = 0; j0 < 32; j0 += 8) {
            if (n0 + j0 >= p.N) break;  // N is a multiple of 8 (checked on the host)
            float o[8];
            float bb[8], rr[8];
            if (EPI == PG_EPI_BIAS || EPI == PG_EPI_BIAS_GELU || EPI == PG_EPI_BIAS_RES)
              unpack<T>(ldg_cached(bias + n0 + j0), bb);
            if (EPI == PG_EPI_BIAS_RES || EPI == PG_EPI_RES) unpack<T>(ldg_cached(R + (size_t)rm * p.ldr + n0 + j0), rr);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float x = v[j0 + j];
              if (EPI == PG_EPI_BIAS || EPI == PG_EPI_BIAS_GELU || EPI == PG_EPI_BIAS_RES) x += bb[j];
              x = rnd<T>(x);
              if (EPI == PG_EPI_BIAS_GELU) x = rnd<T>(gelu_tanh_fast(x));
              if (EPI == PG_EPI_GEGLU) x = rnd<T>(rnd<T>(gelu_tanh_fast(x)) * rnd<T>(u[DUAL ? j0 + j : 0]));
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
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL == 2) {  // the peer may still multicast into this CTA's shared memory or arrive on its barriers
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------ host side
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
    cudaGetLastError();
  }
  return fn;
}

bool make_map_2d_ex(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_cols,
                    int box_rows, int swizzle_bytes, bool is_bf16) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = fn(map, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                  const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

bool make_map_2d(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows,
                 bool is_bf16) {
  return make_map_2d_ex(map, base, rows, cols, ld, 64, box_rows, 128, is_bf16);
}

template <typename T, int EPI, int BN, int CL = 1>
static int launch(const CUtensorMap& ma, const CUtensorMap& mw, const Params& p, cudaStream_t st) {
  constexpr bool DUAL = (EPI == PG_EPI_GEGLU);
  constexpr int NSTAGES = (DUAL || BN > 128) ? 4 : 6;
  constexpr int STAGE_BYTES = TILE_BYTES + (DUAL ? 2 : 1) * BN * BK * 2;
  const size_t smem = 1024 + (size_t)NSTAGES * STAGE_BYTES + 8 * (2 * NSTAGES + 4) + 16;
  auto kern = gemm_tc_kernel<T, EPI, BN, CL>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    set_error("gemm_tc: cannot reserve %zu B of shared memory", smem);
    cudaGetLastError();
    return PG_ERR_CUDA;
  }
  if (CL == 2) {
    const int pairs = cdiv(cdiv(p.M, BM), 2) * cdiv(p.N, BN);
    return launch_tc("gemm_tcgen05_mcast", kern, dim3(2 * (pairs < 74 ? pairs : 74)), dim3(THREADS), smem, 2, pairs <= 148, st, ma, mw, p);
  }
  const int tiles = cdiv(p.M, BM) * cdiv(p.N, BN);
  return launch_tc("gemm_tcgen05", kern, dim3(tiles < 148 ? tiles : 148), dim3(THREADS), smem, 1, tiles <= 296, st, ma, mw, p);
}

}  // namespace tc

bool gemm_tc_supported(int M, int N, int K, int lda, int ldw, int ldc, int epi, int out_f32, int dtype) {
  static const int enabled = env_int("PG_TCGEN05", 1);
  static const int min_m = env_int("PG_TCGEN05_MIN_M", 4);
  if (!enabled || (dtype != PG_BF16 && dtype != PG_F16)) return false;
  if (M < min_m || N % 8 || K % 8 || lda % 8 || ldw % 8 || ldc % (out_f32 ? 4 : 8)) return false;
  if (epi < PG_EPI_NONE || epi > PG_EPI_GEGLU) return false;
  return tc::encode_fn() != nullptr;
}

bool gemm_tc_swap_wanted(int M, int N, int K, int epi, int out_f32, int res_mod);
int gemm_tc_swap(void* C, const void* A, const void* W, const void* R, int M, int N, int K, int lda, int ldw, int ldc,
                 int ldr, int epi, int dtype, cudaStream_t st);
bool gemm_tc_skinny_wanted(int M, int N, int K, int epi);
int gemm_tc_skinny(void* C, const void* A, const void* W, const void* bias, const void* R, int M, int N, int K, int lda,
                   int ldw, int ldc, int ldr, int epi, int out_f32, int dtype, cudaStream_t st);
bool gemm_tc_2cta_wanted(int M, int N, int K, int epi);
int gemm_tc_2cta(void* C, const void* A, const void* W, const void* bias, const void* R, int M, int N, int K, int lda,
                 int ldw, int ldc, int ldr, int res_mod, int epi, int out_f32, int dtype, cudaStream_t st);
int gemm_tc_splitk_factor(int M, int N, int K, int epi, int* bn_out);
int gemm_tc_splitk(void* C, const void* A, const void* W, const void* bias, const void* R, int M, int N, int K, int lda,
                   int ldw, int ldc, int ldr, int res_mod, int epi, int out_f32, int dtype, int bn, int s,
                   cudaStream_t st);

int gemm_tc(void* C, const void* A, const void* W, const void* bias, const void* R, int M, int N, int K, int lda,
            int ldw, int ldc, int ldr, int res_mod, int epi, int out_f32, int dtype, cudaStream_t st) {
  PG_REQUIRE(((uintptr_t)A % 16) == 0 && ((uintptr_t)W % 16) == 0 && ((uintptr_t)C % 16) == 0,
             "gemm_tc: operands must be 16-byte aligned");
  PG_REQUIRE(!R || (((uintptr_t)R % 16) == 0 && ldr % 8 == 0), "gemm_tc: residual must be 16-byte aligned rows");
  PG_REQUIRE(!bias || ((uintptr_t)bias % 16) == 0, "gemm_tc: bias must be 16-byte aligned");
  const bool bf = dtype == PG_BF16;
  if (gemm_tc_swap_wanted(M, N, K, epi, out_f32, res_mod))  // a prompt's worth of rows: CTA pairs, weights as the M operand
    return gemm_tc_swap(C, A, W, R, M, N, K, lda, ldw, ldc, ldr, epi, dtype, st);
  if (res_mod == 0 && gemm_tc_skinny_wanted(M, N, K, epi))  // a few dozen token rows: stream the weights (swap-AB)
    return gemm_tc_skinny(C, A, W, bias, R, M, N, K, lda, ldw, ldc, ldr, epi, out_f32, dtype, st);
  {
    int sk_bn = 0;
    const int s = gemm_tc_splitk_factor(M, N, K, epi, &sk_bn);  // few tiles, long K: a cluster shares each tile
    if (s) return gemm_tc_splitk(C, A, W, bias, R, M, N, K, lda, ldw, ldc, ldr, res_mod, epi, out_f32, dtype, sk_bn, s, st);
  }
  if (gemm_tc_2cta_wanted(M, N, K, epi))  // >= 2 waves of 256 x 256 tiles: CTA pairs (cta_group::2)
    return gemm_tc_2cta(C, A, W, bias, R, M, N, K, lda, ldw, ldc, ldr, res_mod, epi, out_f32, dtype, st);
  // 64-wide N tiles when 128-wide ones would occupy well under one wave of the 148 SMs
  static const int bn_env = env_int("PG_GEMM_BN", 0);
  // and 256-wide ones (half the A-operand shared-memory traffic per FLOP) when there is work for > 2 waves
  int bn = bn_env ? bn_env : ((cdiv(M, tc::BM) * cdiv(N, 128) < 100) ? 64 : (cdiv(M, tc::BM) * cdiv(N, 256) >= 296 ? 256 : 128));
  if (epi == PG_EPI_GEGLU && bn == 256) bn = 128;  // gate+up already fill 512 TMEM columns at 128
  static const int mcast_env = env_int("PG_GEMM_MCAST", 1);
  const bool mcast = mcast_env && bn == 256;
  CUtensorMap ma, mw;
  const int w_rows = (epi == PG_EPI_GEGLU) ? 2 * N : N;
  PG_REQUIRE(tc::make_map_2d(&ma, A, M, K, lda, tc::BM, bf) && tc::make_map_2d(&mw, W, w_rows, K, ldw, mcast ? bn / 2 : bn, bf),
             "gemm_tc: cuTensorMapEncodeTiled failed (M=%d N=%d K=%d lda=%d ldw=%d)", M, N, K, lda, ldw);
  tc::Params p = {C, bias, R, M, N, K, ldc, ldr, res_mod, out_f32};
#define PG_TC(E)                                                                                       \
  if (bn == 64) return bf ? tc::launch<bf16, E, 64>(ma, mw, p, st) : tc::launch<f16, E, 64>(ma, mw, p, st); \
  if (mcast && E != PG_EPI_GEGLU)                                                                        \
    return bf ? tc::launch<bf16, (E == PG_EPI_GEGLU ? PG_EPI_NONE : E), 256, 2>(ma, mw, p, st)           \
              : tc::launch<f16, (E == PG_EPI_GEGLU ? PG_EPI_NONE : E), 256, 2>(ma, mw, p, st);           \
  if (bn == 256 && E != PG_EPI_GEGLU)                                                                    \
    return bf ? tc::launch<bf16, (E == PG_EPI_GEGLU ? PG_EPI_NONE : E), 256>(ma, mw, p, st)              \
              : tc::launch<f16, (E == PG_EPI_GEGLU ? PG_EPI_NONE : E), 256>(ma, mw, p, st);              \
  return bf ? tc::launch<bf16, E, 128>(ma, mw, p, st) : tc::launch<f16, E, 128>(ma, mw, p, st)
  switch (epi) {
    case PG_EPI_NONE: PG_TC(PG_EPI_NONE);
    case PG_EPI_BIAS: PG_TC(PG_EPI_BIAS);
    case PG_EPI_BIAS_GELU: PG_TC(PG_EPI_BIAS_GELU);
    case PG_EPI_BIAS_RES: PG_TC(PG_EPI_BIAS_RES);
    case PG_EPI_RES: PG_TC(PG_EPI_RES);
    case PG_EPI_GEGLU: PG_TC(PG_EPI_GEGLU);
  }
#undef PG_TC
  set_error("gemm_tc: bad epilogue %d", epi);
  return PG_ERR_INVALID;
}

}  // namespace pg
