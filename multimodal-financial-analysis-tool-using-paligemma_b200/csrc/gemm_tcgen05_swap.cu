// Prompt-sized GEMMs (129..512 token rows: the 260-token prefill and the cache-off recompute of the KV ablation):
// C[M,N] = A[M,K] W[N,K]^T with the operands SWAPPED on CTA pairs.
//
// With a few hundred token rows every nn.Linear of the Gemma decoder is a weight-streaming problem (5 GB of weights
// against 1e12 FLOP), but the row-major kernels treat the tokens as the 128-row MMA operand: 260 rows make three
// m-tiles (the third 97 % padding), the weights are re-read per m-tile out of L2 and the L2 -> SM fabric (~6300 B/clk
// chip-wide), not HBM or the tensor pipe, sets the time (gate/up 48 us, down_proj 56 us per layer).  Here:
//   * the WEIGHT rows are the M = 256 operand of a tcgen05.mma.cta_group::2 (128 rows per CTA of the pair), streamed
//     from HBM exactly once;
//   * ALL tokens are the N operand: one accumulator of NT <= 256 columns (<= 256 tokens) or two (<= 512 tokens, the
//     second at TMEM column 256) in TMEM; the pair SPLITS the token tile (each CTA stages NT/2 token rows per k-block), so the token re-read that
//     every weight tile costs is halved against single-CTA swap-AB;
//   * short-N / long-K problems (o_proj, down_proj, q/k/v: 8-10 weight tiles) are split along K over the 74 pairs;
//     every work item leaves an fp32 partial [split][token][feature] in a caller-provided workspace (L2-resident) and
//     gemm_swap_reduce_kernel sums the splits and applies the epilogue (residual / GeGLU / rounding points of
//     gemm_simt.cu) with fully coalesced rows.
// Roles per CTA (as gemm_tcgen05_2cta.cu): warp 0 TMA producer, warp 1 (leader CTA) MMA issuer, warp 2 TMEM
// allocation, warps 4-11 epilogue (TMEM -> fp32 partial rows).
#include "tc_common.cuh"

namespace pg {
namespace tc {

constexpr int SW_BK = 64, SW_WROWS = 128, SW_THREADS = 384;
constexpr int SW_W_BYTES = SW_WROWS * SW_BK * 2;   // 16 KB of weights per CTA per k-block
constexpr int SW_SMEM_BUDGET = 220 * 1024;

struct ParamsSw {
  float* P;            // workspace: [splits][M tokens][N features] fp32
  int M, N, K;         // tokens, output features, reduction length
  int NT, n_t;         // token tile width (multiple of 16, <= 256) and number of tiles (1 or 2)
  int splits, kb_per_split, stages;
  int dbg;             // PG_SWAP_DBG bit 0: skip the partial stores (timing experiments only)
};

__device__ __forceinline__ uint32_t sw_mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void sw_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void sw_tma_cg2(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void sw_umma_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void sw_commit_cg2(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void sw_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int FMT>  // 1 bf16, 0 f16
__global__ void __launch_bounds__(SW_THREADS, 1)
gemm_swap_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x, ParamsSw p) {
  constexpr int MAX_STAGES = 8;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int x_bytes = (p.NT / 2) * SW_BK * 2;                 // one token tile's half per k-block (multiple of 1024)
  const int stage_bytes = SW_W_BYTES + p.n_t * x_bytes;
  const uint32_t bars = smem_base + p.stages * stage_bytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };                      // used in the leader CTA
  auto empty_bar = [&](int s) { return bars + 8u * (MAX_STAGES + s); };      // one per CTA
  const uint32_t tfull_bar = bars + 8u * (2 * MAX_STAGES);                   // one per CTA
  const uint32_t tempty_bar = bars + 8u * (2 * MAX_STAGES + 1);              // used in the leader CTA
  const uint32_t tmem_slot = bars + 8u * (2 * MAX_STAGES + 2);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t crank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const bool leader = crank == 0;
  const int f_tiles = (p.N + 2 * SW_WROWS - 1) / (2 * SW_WROWS);
  const int items = f_tiles * p.splits;
  const int first = (int)(blockIdx.x / 2), stride = (int)(gridDim.x / 2);
  const int k_blocks = (p.K + SW_BK - 1) / SW_BK;
  const uint32_t idesc = umma_idesc(FMT, 2 * SW_WROWS, p.NT);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_x)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 16);  // 8 epilogue warps x 2 CTAs
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  sw_cluster_sync();  // the peer's barriers and TMEM exist before anything lands on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs): own 128 weight rows + own half of every token tile =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      bool waited = false;
      for (int item = first; item < items; item += stride) {
        const int ft = item / p.splits, sp = item % p.splits;
        const int kb0 = sp * p.kb_per_split, kb1 = min(k_blocks, kb0 + p.kb_per_split);
        // every pair walks its k-blocks from a different starting point: at any moment the pairs read different
        // token slices, so the few hundred L2 lines of one slice are not requested by all 148 SMs at once
        const int nkb = kb1 - kb0, rot = (p.dbg & 2) ? 0 : (first * 5) % nkb;
        for (int i = 0; i < nkb; ++i) {
          const int kb = kb0 + (i + rot) % nkb;
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * stage_bytes;
          const uint32_t full_leader = sw_mapa(full_bar(stage), 0);
          if (leader) mbar_expect_tx(full_bar(stage), 2 * stage_bytes);  // both CTAs' bytes
          // the weights never depend on the preceding kernels: their loads go out before the dependency wait
          sw_tma_cg2(sa, &map_w, full_leader, kb * SW_BK, ft * 2 * SW_WROWS + (int)crank * SW_WROWS);
          if (!waited) { pdl_wait(); waited = true; }   // the token rows come from the predecessor
          for (int t = 0; t < p.n_t; ++t)
            sw_tma_cg2(sa + SW_W_BYTES + t * x_bytes, &map_x, full_leader, kb * SW_BK, t * p.NT + (int)crank * (p.NT / 2));
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
      if (!waited) pdl_wait();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && lane == 0) {
      int stage = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int item = first; item < items; item += stride) {
        const int sp = item % p.splits;
        const int kb0 = sp * p.kb_per_split, kb1 = min(k_blocks, kb0 + p.kb_per_split);
        mbar_wait(tempty_bar, acc_phase ^ 1);  // both CTAs' epilogues have drained the accumulator
        tc_fence_after();
        for (int i = kb0; i < kb1; ++i) {      // the producer's k order (rotated per pair); the sum does not care
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * stage_bytes;
#pragma unroll
          for (int k = 0; k < SW_BK / 16; ++k)
            for (int t = 0; t < p.n_t; ++t)
              sw_umma_cg2(tmem_base + t * 256, umma_desc(sa + k * 32), umma_desc(sa + SW_W_BYTES + t * x_bytes + k * 32), idesc,
                          (i > kb0 || k > 0) ? 1u : 0u);
          sw_commit_cg2(empty_bar(stage), (uint16_t)0x3);  // the stage is free in BOTH CTAs once these MMAs have read it
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        sw_commit_cg2(tfull_bar, (uint16_t)0x3);           // accumulator complete, in both CTAs' TMEM
        acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM -> fp32 partial [split][token][feature] (both CTAs) =====================
    const int q = warp & 3, half = (warp - 4) >> 2;  // TMEM lane quarter; the two halves interleave 16-column chunks
    uint32_t acc_phase = 0;
    const int chunks_per_tile = p.NT / 16, chunks = p.n_t * chunks_per_tile;
    pdl_wait();  // the workspace may still be read by the predecessor's reduce kernel
    for (int item = first; item < items; item += stride) {
      const int ft = item / p.splits, sp = item % p.splits;
      const int f = ft * 2 * SW_WROWS + (int)crank * SW_WROWS + q * 32 + lane;   // this thread's output feature
      float* prow = p.P + (size_t)sp * p.M * p.N + f;
      mbar_wait(tfull_bar, acc_phase);
      tc_fence_after();
      const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
      for (int c = half; c < chunks; c += 2) {
        const int t = c / chunks_per_tile, cc = c % chunks_per_tile;   // token tile t lives at TMEM columns [256 t, 256 t + NT)
        float v[16];
        tmem_ld16(t_lane + t * 256 + cc * 16, v);
        tmem_ld_wait();
        if (f < p.N && !(p.dbg & 1)) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int tok = t * p.NT + cc * 16 + j;
            if (tok < p.M) prow[(size_t)tok * p.N] = v[j];   // a warp writes 32 consecutive features: one 128-byte row piece
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) sw_arrive_cluster(sw_mapa(tempty_bar, 0));
      acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  sw_cluster_sync();  // the leader's MMAs read this CTA's shared memory, its commits arrive on this CTA's barriers
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// out[m][n] = epilogue(sum over splits of P[s][m][n]); rounding points of gemm_simt.cu / gemm_tc_kernel.
// GeGLU: W = [Wgate; Wup], P has 2*Nout columns, out[m][f] = rnd(rnd(gelu(rnd(g))) * rnd(u)).
template <typename T, int EPI>
__global__ void __launch_bounds__(256)
gemm_swap_reduce_kernel(T* __restrict__ out, const float* __restrict__ P, const T* __restrict__ R, int M, int Nout, int Np,
                        int splits, int ldc, int ldr) {
  pdl_launch_dependents();
  pdl_wait();
  const int vec_per_row = Nout / 4;
  const long long total = (long long)M * vec_per_row;
  const size_t split_stride = (size_t)M * Np;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(i / vec_per_row), n = (int)(i % vec_per_row) * 4;
    const float* src = P + (size_t)m * Np + n;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), u = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < splits; ++s) {
      const float4 v = *reinterpret_cast<const float4*>(src + s * split_stride);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      if (EPI == PG_EPI_GEGLU) {
        const float4 w = *reinterpret_cast<const float4*>(src + s * split_stride + Nout);
        u.x += w.x; u.y += w.y; u.z += w.z; u.w += w.w;
      }
    }
    float o[4] = {rnd<T>(a.x), rnd<T>(a.y), rnd<T>(a.z), rnd<T>(a.w)};
    if (EPI == PG_EPI_GEGLU) {
      const float uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = rnd<T>(rnd<T>(gelu_tanh_fast(o[j])) * rnd<T>(uu[j]));
    }
    if (EPI == PG_EPI_RES) {
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = rnd<T>(o[j] + to_f<T>(R[(size_t)m * ldr + n + j]));
    }
    T* dst = out + (size_t)m * ldc + n;
#pragma unroll
    for (int j = 0; j < 4; ++j) dst[j] = from_f<T>(o[j]);
  }
}

}  // namespace tc

// ---------------------------------------------------------------------------------------- host side
static void* g_ws_ptr = nullptr;
static long long g_ws_bytes = 0;
static thread_local bool g_force_swap = false;   // pg_gemm(impl = 3): take this kernel for every epilogue it supports

void gemm_tc_swap_force(bool on) { g_force_swap = on; }

extern "C" int pg_set_workspace(void* ptr, long long bytes) {
  g_ws_ptr = ptr;
  g_ws_bytes = ptr ? bytes : 0;
  return PG_OK;
}

static void swap_plan(int M, int N, int K, tc::ParamsSw* p) {
  p->M = M; p->N = N; p->K = K;
  p->n_t = M <= 256 ? 1 : 2;
  p->NT = (((M + p->n_t - 1) / p->n_t) + 15) / 16 * 16;
  const int f_tiles = cdiv(N, 2 * tc::SW_WROWS), k_blocks = cdiv(K, tc::SW_BK);
  // split K until the work items fill the 74 CTA pairs about once (every split keeps >= 2 k-blocks)
  int splits = 1;
  if (f_tiles < 74) {
    splits = 74 / f_tiles;
    if (splits > k_blocks / 2) splits = k_blocks / 2 > 0 ? k_blocks / 2 : 1;
  }
  int kbs = cdiv(k_blocks, splits);
  splits = cdiv(k_blocks, kbs);           // no empty split
  p->splits = splits;
  p->kb_per_split = kbs;
  const int stage_bytes = tc::SW_W_BYTES + p->n_t * (p->NT / 2) * tc::SW_BK * 2;
  int stages = (tc::SW_SMEM_BUDGET - 1024 - 256) / stage_bytes;
  p->stages = stages > 8 ? 8 : stages;
  static const int dbg = env_int("PG_SWAP_DBG", 0), max_stages = env_int("PG_SWAP_STAGES", 8);
  p->dbg = dbg;
  if (p->stages > max_stages) p->stages = max_stages;
}

bool gemm_tc_swap_wanted(int M, int N, int K, int epi, int out_f32, int res_mod) {
  static const int enabled = env_int("PG_GEMM_SWAP", 1);
  static const int min_m = env_int("PG_GEMM_SWAP_MIN_M", 4);
  if ((!enabled && !g_force_swap) || out_f32 || res_mod != 0 || M < min_m || M > 512) return false;
  if (epi != PG_EPI_NONE && epi != PG_EPI_RES && epi != PG_EPI_GEGLU) return false;
  // which projections take this kernel (measured per projection, profiles/README.md): bit 0 q/k/v (no epilogue),
  // bit 1 o_proj (residual, K < 8192), bit 2 down_proj (residual, K >= 8192), bit 3 gate/up (GeGLU)
  // measured: prompt-sized rows (129..512) -- down_proj only (prefill 4.16 -> 3.45 ms; q/k/v, o_proj and gate/up are
  // faster on the row-major kernels); batched-decode rows (4..128) -- o_proj and down_proj (batch 32: 1.93 -> 1.86 ms)
  static const int mask_large = env_int("PG_SWAP_MASK", 4), mask_small = env_int("PG_SWAP_MASK_SMALL", 6);
  const int mask = M > 128 ? mask_large : mask_small;
  const int kind = epi == PG_EPI_NONE ? 1 : (epi == PG_EPI_GEGLU ? 8 : (K >= 8192 ? 4 : 2));
  if (!(mask & kind) && !g_force_swap) return false;
  const int n_w = epi == PG_EPI_GEGLU ? 2 * N : N;   // weight rows
  if (n_w % 4 || N % 4) return false;
  tc::ParamsSw p;
  swap_plan(M, n_w, K, &p);
  if (p.stages < 3) return false;
  return g_ws_ptr && (long long)p.splits * M * n_w * 4 <= g_ws_bytes;
}

int gemm_tc_swap(void* C, const void* A, const void* W, const void* R, int M, int N, int K, int lda, int ldw, int ldc,
                 int ldr, int epi, int dtype, cudaStream_t st) {
  const bool bf = dtype == PG_BF16;
  const int n_w = epi == PG_EPI_GEGLU ? 2 * N : N;
  tc::ParamsSw p;
  swap_plan(M, n_w, K, &p);
  p.P = reinterpret_cast<float*>(g_ws_ptr);
  CUtensorMap mw, mx;
  PG_REQUIRE(tc::make_map_2d(&mw, W, n_w, K, ldw, tc::SW_WROWS, bf) && tc::make_map_2d(&mx, A, M, K, lda, p.NT / 2, bf),
             "gemm_swap: cuTensorMapEncodeTiled failed (M=%d N=%d K=%d lda=%d ldw=%d)", M, n_w, K, lda, ldw);
  const int stage_bytes = tc::SW_W_BYTES + p.n_t * (p.NT / 2) * tc::SW_BK * 2;
  const size_t smem = 1024 + (size_t)p.stages * stage_bytes + 8 * (2 * 8 + 3) + 16;
  auto kern = bf ? tc::gemm_swap_kernel<1> : tc::gemm_swap_kernel<0>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    set_error("gemm_swap: cannot reserve %zu B of shared memory", smem);
    cudaGetLastError();
    return PG_ERR_CUDA;
  }
  const int items = cdiv(n_w, 2 * tc::SW_WROWS) * p.splits;
  int rc = launch_tc("gemm_tcgen05_swap", kern, dim3(2 * (items < 74 ? items : 74)), dim3(tc::SW_THREADS), smem, 2, true, st,
                     mw, mx, p);
  if (rc != PG_OK) return rc;
  const long long vecs = (long long)M * (N / 4);
  const int grid = (int)((vecs + 255) / 256 < 148 * 8 ? (vecs + 255) / 256 : 148 * 8);
#define PG_SWR(T_, E_)                                                                                          \
  return launch_tc("gemm_swap_reduce", tc::gemm_swap_reduce_kernel<T_, E_>, dim3(grid), dim3(256), 0, 1, true, st, \
                   (T_*)C, (const float*)p.P, (const T_*)R, M, N, n_w, p.splits, ldc, ldr)
  if (bf) {
    if (epi == PG_EPI_NONE) { PG_SWR(bf16, PG_EPI_NONE); }
    if (epi == PG_EPI_RES) { PG_SWR(bf16, PG_EPI_RES); }
    PG_SWR(bf16, PG_EPI_GEGLU);
  } else {
    if (epi == PG_EPI_NONE) { PG_SWR(f16, PG_EPI_NONE); }
    if (epi == PG_EPI_RES) { PG_SWR(f16, PG_EPI_RES); }
    PG_SWR(f16, PG_EPI_GEGLU);
  }
#undef PG_SWR
}

}  // namespace pg
