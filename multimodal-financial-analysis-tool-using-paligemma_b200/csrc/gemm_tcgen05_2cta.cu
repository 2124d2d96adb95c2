// C[M,N] = A[M,K] W[N,K]^T with CTA pairs: tcgen05.mma.cta_group::2, one 256 x 256 tile per 2-CTA cluster.
//
// The single-CTA kernel (gemm_tcgen05.cu) tops out near half the tensor peak: an M=128 cta_group::1 MMA
// dispatches at max(M,128)*N/256 cycles, and every SM reads the full W tile from its own shared memory.
// Here the two SMs of a TPC share one MMA: each CTA stages ITS 128 rows of A and ITS 128 of the 256 W rows
// (32 KB per k-block instead of 48), the leader CTA's elected thread issues M=256 x N=256 x K=16 MMAs that
// read both CTAs' shared memory, and each CTA's TMEM receives its own 128 x 256 fp32 accumulator.
//
//   warp 0 (both CTAs)   TMA producer: own A tile + own W half; completion bytes land on the LEADER's
//                        full barrier (cp.async.bulk.tensor ... .cta_group::2)
//   warp 1 (leader only) MMA issuer; tcgen05.commit.cta_group::2 multicast frees the stage in both CTAs
//                        and publishes the accumulator to both epilogues
//   warp 2 (both CTAs)   TMEM allocation (tcgen05.alloc.cta_group::2, 512 columns = 2 accumulator stages)
//   warps 4-11           epilogue: two warps per TMEM lane quarter, 128 columns each; arrive on the leader's
//                        tmem-empty barrier (local or remote)
// Same fused epilogues and rounding points as gemm_tc_kernel (GeGLU stays on the single-CTA kernel).
#include "tc_common.cuh"

namespace pg {
namespace tc {

constexpr int G2_BM = 128, G2_BN = 256, G2_BK = 64;
constexpr int G2_STAGES = 6, G2_THREADS = 384;
constexpr int G2_A_BYTES = G2_BM * G2_BK * 2, G2_W_BYTES = (G2_BN / 2) * G2_BK * 2;
constexpr int G2_STAGE_BYTES = G2_A_BYTES + G2_W_BYTES;  // 32 KB per CTA per k-block
constexpr int G2_EPI_WARPS = 8, G2_STG_BYTES = 32 * 128;  // per-warp epilogue staging: 32 rows x 64 16-bit columns

struct Params2 {
  void* C;
  const void* bias;
  const void* R;
  int M, N, K, ldc, ldr, res_mod, out_f32;
  int m_major;  // tile order: consecutive tiles share the A rows (W small enough to stay in L2) or the W rows
};

__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion bytes are signalled on an mbarrier that may live in the peer CTA of the pair
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_cg2(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// one output element: bias, rounding to T, GELU, residual -- the rounding points of gemm_tc_kernel / gemm_simt.cu
template <typename T, int EPI>
__device__ __forceinline__ float epi_apply(float x, float b, float r) {
  if (EPI == PG_EPI_BIAS || EPI == PG_EPI_BIAS_GELU || EPI == PG_EPI_BIAS_RES) x += b;
  x = rnd<T>(x);
  if (EPI == PG_EPI_BIAS_GELU) x = rnd<T>(gelu_tanh_fast(x));
  if (EPI == PG_EPI_BIAS_RES || EPI == PG_EPI_RES) x = rnd<T>(x + r);
  return x;
}
// two elements at once with packed conversions (keeps the XU pipe for MUFU.TANH)
template <typename T, int EPI>
__device__ __forceinline__ void epi_apply2(float& x0, float& x1, float b0, float b1, float r0, float r1) {
  if (EPI == PG_EPI_BIAS || EPI == PG_EPI_BIAS_GELU || EPI == PG_EPI_BIAS_RES) { x0 += b0; x1 += b1; }
  rnd2<T>(x0, x1);
  if (EPI == PG_EPI_BIAS_GELU) { x0 = gelu_tanh_fast(x0); x1 = gelu_tanh_fast(x1); rnd2<T>(x0, x1); }
  if (EPI == PG_EPI_BIAS_RES || EPI == PG_EPI_RES) { x0 += r0; x1 += r1; rnd2<T>(x0, x1); }
}

template <typename T, int EPI>
__global__ void __launch_bounds__(G2_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, Params2 p) {
  constexpr uint32_t IDESC = umma_idesc(std::is_same<T, bf16>::value ? 1 : 0, 2 * G2_BM, G2_BN);
  // a last n-tile with <= 128 real columns (N = 1152: o_proj, fc2) runs N = 128 MMAs: each CTA supplies 64 W rows
  constexpr uint32_t IDESC_HALF = umma_idesc(std::is_same<T, bf16>::value ? 1 : 0, 2 * G2_BM, G2_BN / 2);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stg_base = smem_base + G2_STAGES * G2_STAGE_BYTES;
  const uint32_t bars = stg_base + G2_EPI_WARPS * G2_STG_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };                       // used in the leader CTA
  auto empty_bar = [&](int s) { return bars + 8u * (G2_STAGES + s); };        // one per CTA
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * G2_STAGES + a); };    // one per CTA
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * G2_STAGES + 2 + a); };  // used in the leader CTA
  const uint32_t tmem_slot = bars + 8u * (2 * G2_STAGES + 4);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t crank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const bool leader = crank == 0;
  const int m_pairs = (((p.M + G2_BM - 1) / G2_BM) + 1) / 2, n_tiles = (p.N + G2_BN - 1) / G2_BN;
  const int total_tiles = m_pairs * n_tiles;
  const int sched_first = (int)(blockIdx.x / 2), sched_stride = (int)(gridDim.x / 2);
  const int k_blocks = (p.K + G2_BK - 1) / G2_BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < G2_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 16); }  // 8 epilogue warps x 2 CTAs
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers and TMEM exist before anything lands on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      pdl_wait();  // A (and everything else) comes from the preceding kernels
      for (int tile = sched_first; tile < total_tiles; tile += sched_stride) {
        const int n_blk = p.m_major ? tile % n_tiles : tile / m_pairs;  // neighbouring clusters share A rows or W rows (L2 reuse)
        const int m_blk = 2 * (p.m_major ? tile / n_tiles : tile % m_pairs) + (int)crank;  // a ghost tile past M loads zeros, stores nothing
        const int w_step = (n_blk * G2_BN + G2_BN / 2 >= p.N) ? G2_BN / 4 : G2_BN / 2;  // half tile: 64 rows per CTA
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * G2_STAGE_BYTES;
          const uint32_t full_leader = mapa_u32(full_bar(stage), 0);
          if (leader) mbar_expect_tx(full_bar(stage), 2 * G2_STAGE_BYTES);  // both CTAs' bytes
          tma_load_2d_cg2(sa, &map_a, full_leader, kb * G2_BK, m_blk * G2_BM);
          tma_load_2d_cg2(sa + G2_A_BYTES, &map_w, full_leader, kb * G2_BK, n_blk * G2_BN + (int)crank * w_step);
          if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && lane == 0) {
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = sched_first; tile < total_tiles; tile += sched_stride) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);  // both CTAs' epilogues have drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * G2_BN;
        const int n_blk_mma = p.m_major ? tile % n_tiles : tile / m_pairs;
        const uint32_t idesc = (n_blk_mma * G2_BN + G2_BN / 2 >= p.N) ? IDESC_HALF : IDESC;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * G2_STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < G2_BK / 16; ++k)
            umma_cg2(d_tmem, umma_desc(sa + k * 32), umma_desc(sa + G2_A_BYTES + k * 32), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit_cg2(empty_bar(stage), (uint16_t)0x3);  // the stage is free in BOTH CTAs once these MMAs have read it
          if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_cg2(tfull_bar(acc), (uint16_t)0x3);       // accumulator complete, in both CTAs' TMEM
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM -> registers -> global (both CTAs) =====================
    const int q = warp & 3, half = (warp - 4) >> 2;  // TMEM lane quarter, column half
    int acc = 0;
    uint32_t acc_phase = 0;
    T* Ct = reinterpret_cast<T*>(p.C);
    float* Cf = reinterpret_cast<float*>(p.C);
    const T* bias = reinterpret_cast<const T*>(p.bias);
    const T* R = reinterpret_cast<const T*>(p.R);
    constexpr bool HAS_BIAS = EPI == PG_EPI_BIAS || EPI == PG_EPI_BIAS_GELU || EPI == PG_EPI_BIAS_RES;
    constexpr bool HAS_RES = EPI == PG_EPI_BIAS_RES || EPI == PG_EPI_RES;
    uint8_t* wbuf = smem_raw + (stg_base - smem_u32(smem_raw)) + (warp - 4) * G2_STG_BYTES;
    pdl_wait();  // residual reads and output writes only after the predecessor has finished
    for (int tile = sched_first; tile < total_tiles; tile += sched_stride) {
      const int n_blk = p.m_major ? tile % n_tiles : tile / m_pairs;
      const int m_blk = 2 * (p.m_major ? tile / n_tiles : tile % m_pairs) + (int)crank;
      const int row0 = m_blk * G2_BM + q * 32;  // first row of this warp's TMEM lane quarter
      const int m = row0 + lane;
      const int ch_l = lane & 7, r_l = lane >> 3;  // staging <-> global mapping: 8 lanes per 128-byte row segment
      const int nw0 = n_blk * G2_BN + half * (G2_BN / 2);
      // residual rows of both 64-column slices are requested BEFORE waiting for the accumulator: ordinary loads
      // queue behind ~190 KB of in-flight TMA data per SM, and this wait is where the epilogue warps idle anyway
      uint4 r1[8];
      if (HAS_RES && !p.out_f32) {
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int rr_ = it * 4 + r_l, mm = row0 + rr_, col = nw0 + ch_l * 8;
          const T* rrow = R + (size_t)(p.res_mod > 0 ? (mm % p.res_mod) : mm) * p.ldr;
          uint4 u0 = make_uint4(0, 0, 0, 0);
          r1[it] = make_uint4(0, 0, 0, 0);
          if (mm < p.M && col < p.N) u0 = ldg_cached(rrow + col);
          if (mm < p.M && col + 64 < p.N) r1[it] = ldg_cached(rrow + col + 64);
          *reinterpret_cast<uint4*>(wbuf + rr_ * 128 + ((ch_l ^ (rr_ & 7)) << 4)) = u0;
        }
        __syncwarp();
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * G2_BN + half * (G2_BN / 2);
      if (p.out_f32) {
        // fp32 rows (logits): direct 32-byte stores, thread <-> row
#pragma unroll 1
        for (int c = 0; c < G2_BN / 64; ++c) {
          float v[32];
          tmem_ld32(t_row + c * 32, v);
          tmem_ld_wait();
          const int n0 = n_blk * G2_BN + half * (G2_BN / 2) + c * 32;
          if (m < p.M && n0 < p.N) {
            const int rm = p.res_mod > 0 ? (m % p.res_mod) : m;
#pragma unroll
            for (int j0 = 0; j0 < 32; j0 += 8) {
              if (n0 + j0 >= p.N) break;  // N is a multiple of 8 (checked on the host)
              float o[8], bb[8], rr[8];
              if (HAS_BIAS) unpack<T>(ldg_cached(bias + n0 + j0), bb);
              if (HAS_RES) unpack<T>(ldg_cached(R + (size_t)rm * p.ldr + n0 + j0), rr);
#pragma unroll
              for (int j = 0; j < 8; ++j) o[j] = epi_apply<T, EPI>(v[j0 + j], bb[j], rr[j]);
              float4* dst = reinterpret_cast<float4*>(Cf + (size_t)m * p.ldc + n0 + j0);
              dst[0] = make_float4(o[0], o[1], o[2], o[3]);
              dst[1] = make_float4(o[4], o[5], o[6], o[7]);
            }
          }
        }
      } else {
        // 16-bit rows: 64-column slices through a per-warp staging tile so that the residual loads and the output
        // stores are whole 128-byte row segments (4 rows per instruction) instead of 32 scattered 16-byte pieces
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
          const int n0 = nw0 + sl * 64;
          if (n0 >= p.N || row0 >= p.M) break;  // warp-uniform
          if (HAS_RES && sl == 1) {
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int rr_ = it * 4 + r_l;
              *reinterpret_cast<uint4*>(wbuf + rr_ * 128 + ((ch_l ^ (rr_ & 7)) << 4)) = r1[it];
            }
            __syncwarp();
          }
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            float v[32];
            tmem_ld32(t_row + sl * 64 + cc * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int j0 = 0; j0 < 32; j0 += 8) {
              const int ch = cc * 4 + j0 / 8;
              uint4* slot = reinterpret_cast<uint4*>(wbuf + lane * 128 + ((ch ^ (lane & 7)) << 4));
              float o[8], bb[8], rr[8];
              if (HAS_BIAS) {
                if (n0 + ch * 8 < p.N) unpack<T>(ldg_cached(bias + n0 + ch * 8), bb);
                else {
#pragma unroll
                  for (int j = 0; j < 8; ++j) bb[j] = 0.f;
                }
              }
              if (HAS_RES) unpack<T>(*slot, rr);
#pragma unroll
              for (int j = 0; j < 8; j += 2) {
                o[j] = v[j0 + j]; o[j + 1] = v[j0 + j + 1];
                epi_apply2<T, EPI>(o[j], o[j + 1], bb[j], bb[j + 1], rr[j], rr[j + 1]);
              }
              *slot = pack<T>(o);
            }
          }
          __syncwarp();
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr_ = it * 4 + r_l, mm = row0 + rr_, col = n0 + ch_l * 8;
            if (mm < p.M && col < p.N)
              *reinterpret_cast<uint4*>(Ct + (size_t)mm * p.ldc + col) =
                  *reinterpret_cast<const uint4*>(wbuf + rr_ * 128 + ((ch_l ^ (rr_ & 7)) << 4));
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(tempty_bar(acc), 0));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the leader's MMAs read this CTA's shared memory, its commits arrive on this CTA's barriers
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

template <typename T, int EPI>
static int launch2(const CUtensorMap& ma, const CUtensorMap& mw, const Params2& p, cudaStream_t st) {
  const size_t smem = 1024 + (size_t)G2_STAGES * G2_STAGE_BYTES + G2_EPI_WARPS * G2_STG_BYTES + 8 * (2 * G2_STAGES + 4) + 16;
  auto kern = gemm_tc2_kernel<T, EPI>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    set_error("gemm_tc2: cannot reserve %zu B of shared memory", smem);
    cudaGetLastError();
    return PG_ERR_CUDA;
  }
  const int pairs = cdiv(cdiv(p.M, G2_BM), 2) * cdiv(p.N, G2_BN);
  return launch_tc("gemm_tcgen05_2cta", kern, dim3(2 * (pairs < 74 ? pairs : 74)), dim3(G2_THREADS), smem, 2, pairs <= 148, st, ma, mw, p);
}

}  // namespace tc

// CTA-pair GEMM: taken for problems with at least ~2 waves of 256 x 256 tiles (PG_GEMM_2CTA=0 disables it).
bool gemm_tc_2cta_wanted(int M, int N, int K, int epi) {
  static const int enabled = env_int("PG_GEMM_2CTA", 1);
  static const int min_pairs = env_int("PG_GEMM_2CTA_MIN_PAIRS", 148);
  if (!enabled || epi == PG_EPI_GEGLU) return false;
  return cdiv(cdiv(M, tc::G2_BM), 2) * cdiv(N, tc::G2_BN) >= min_pairs;
}

int gemm_tc_2cta(void* C, const void* A, const void* W, const void* bias, const void* R, int M, int N, int K, int lda,
                 int ldw, int ldc, int ldr, int res_mod, int epi, int out_f32, int dtype, cudaStream_t st) {
  const bool bf = dtype == PG_BF16;
  CUtensorMap ma, mw;
  PG_REQUIRE(tc::make_map_2d(&ma, A, M, K, lda, tc::G2_BM, bf) && tc::make_map_2d(&mw, W, N, K, ldw, tc::G2_BN / 2, bf),
             "gemm_tc2: cuTensorMapEncodeTiled failed (M=%d N=%d K=%d lda=%d ldw=%d)", M, N, K, lda, ldw);
  // A is re-read once per n-tile (n-major order) or W once per m-pair (m-major): stream the big operand once and
  // keep the small one in L2
  static const int order_env = env_int("PG_GEMM_2CTA_ORDER", -1);
  const int m_major = order_env >= 0 ? order_env : ((long long)N * K * 2 <= (32ll << 20) ? 1 : 0);
  tc::Params2 p = {C, bias, R, M, N, K, ldc, ldr, res_mod, out_f32, m_major};
#define PG_TC2(E) return bf ? tc::launch2<bf16, E>(ma, mw, p, st) : tc::launch2<f16, E>(ma, mw, p, st)
  switch (epi) {
    case PG_EPI_NONE: PG_TC2(PG_EPI_NONE);
    case PG_EPI_BIAS: PG_TC2(PG_EPI_BIAS);
    case PG_EPI_BIAS_GELU: PG_TC2(PG_EPI_BIAS_GELU);
    case PG_EPI_BIAS_RES: PG_TC2(PG_EPI_BIAS_RES);
    case PG_EPI_RES: PG_TC2(PG_EPI_RES);
  }
#undef PG_TC2
  set_error("gemm_tc2: bad epilogue %d", epi);
  return PG_ERR_INVALID;
}

}  // namespace pg
