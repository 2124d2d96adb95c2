// Unmasked softmax attention on the tcgen05 tensor cores (prefill, cache-off recompute, SigLIP).
// Reference: GemmaAttention.forward (modeling_gemma.py:262-288, additive mask all zeros) and
// SiglipAttention.forward (modeling_siglip.py:116-131); fp32 softmax, scale applied after QK^T.
//
// One CTA = 128 query rows of one (batch, head).  Two passes over the keys, both on tensor cores:
//   pass 0: S = Q K^T per key tile (tcgen05.mma, fp32 in TMEM) -> each thread owns one query row,
//           reads its S row with tcgen05.ld and keeps the running max (no shuffles at all);
//   pass 1: S again, p = exp(s - max) -> bf16 P tile written to shared memory in the K-major
//           SWIZZLE_128B layout -> O += P V with V consumed as an MN-major operand straight from its
//           natural [key, dim] layout (no transpose), O accumulating in TMEM with no rescaling.
// Recomputing QK^T costs tensor-core time that is idle anyway and removes the online-softmax
// correction of O.  Q, K, V tiles arrive by TMA; K/V may be contiguous or paged (page == key tile).
// Head rows may be padded (hd_stride > hd): the SigLIP path pads 72 -> 128 with zero weights.
#include <cmath>

#include "tc_common.cuh"

namespace pg {
namespace tc {

constexpr int AQ = 128;  // query rows per CTA == TMEM lanes

__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  // MN-major, SWIZZLE_128B: 64 contiguous MN elements (128 B) per K row, 8-row groups 1024 B apart (SBO),
  // next 64-element MN block lbo_bytes away (LBO)
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | (64ull << 32) |
         (1ull << 46) | (2ull << 61);
}

struct AttnParams {
  void* out;
  int ld_out, hd, q_len, n_heads, kv_group;      // kv_group = n_heads / n_kv_heads
  int q_col0, k_col0, v_col0, hd_stride;         // column of head 0 in the Q / K / V tensors, distance between heads
  long long kv_batch_rows;                       // contiguous K/V: rows per batch element
  const int32_t* page_table;                     // paged K/V (page_size == key tile) or NULL
  int pt_stride;
  const int32_t* kv_len_dev;
  int kv_len_const, kv_len_add;
  float scale;
  int scale_mode;                                // 0: s*scale, 1: s/scale
  float scale_mul;                               // != 0: the scaling is exactly s*scale_mul (power-of-two divisor)
};

template <typename T, int HDP, int KT>  // HDP: padded head dim in shared memory (128 or 256); KT: keys per tile
__global__ void __launch_bounds__(256, HDP == 128 ? 2 : 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                    const __grid_constant__ CUtensorMap map_v, AttnParams p) {
  constexpr int KB = HDP / 64;                   // 64-column blocks of the head dim
  constexpr int Q_BYTES = AQ * HDP * 2, K_BYTES = KT * HDP * 2, V_BYTES = K_BYTES, P_BYTES = AQ * KT * 2;
  constexpr int S_COL = 0, O_COL = KT;           // TMEM columns
  constexpr int TMEM_NEED = KT + HDP, TMEM_ALLOC = TMEM_NEED <= 256 ? 256 : 512;  // 256 lets two CTAs share an SM
  constexpr int FMT = std::is_same<T, bf16>::value ? 1 : 0;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // K/V tiles are double-buffered: tile it+1 is requested as soon as S(it) has been issued, so its TMA latency hides
  // behind the softmax of tile it instead of heading every step of the chain
  const uint32_t s_q = base, s_kv = s_q + Q_BYTES, s_p = s_kv + 2 * (K_BYTES + V_BYTES);
  auto s_k = [&](int bf) { return s_kv + bf * (K_BYTES + V_BYTES); };
  auto s_v = [&](int bf) { return s_kv + bf * (K_BYTES + V_BYTES) + K_BYTES; };
  const uint32_t bars = s_p + P_BYTES;
  const uint32_t bar_q = bars, bar_kv0 = bars + 8, bar_s = bars + 16, bar_sfree = bars + 24, bar_p = bars + 32,
                 bar_o = bars + 40, bar_kv1 = bars + 48, tmem_slot = bars + 56;
  auto bar_kv = [&](int bf) { return bf ? bar_kv1 : bar_kv0; };
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  uint8_t* p_ptr = smem_raw + (s_p - smem_u32(smem_raw));

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z, kvh = h / p.kv_group;
  const int nv = ((p.hd + 15) / 16) * 16;        // PV output columns (multiple of 16)
  const int k_steps = (p.hd + 15) / 16;          // 16-wide K steps of QK^T that hold real data

  if (warp == 1 && lane == 0) {
    mbar_init(bar_q, 1); mbar_init(bar_kv0, 1); mbar_init(bar_kv1, 1); mbar_init(bar_s, 1); mbar_init(bar_sfree, 128);
    mbar_init(bar_p, 128); mbar_init(bar_o, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_ALLOC) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();  // q / k / v, lengths and page tables come from the preceding kernels
  const int T_len = p.kv_len_dev ? (p.kv_len_dev[b] + p.kv_len_add) : p.kv_len_const;
  const int n_tiles = (T_len + KT - 1) / KT;

  if (warp == 0 && lane == 0) {
    // ===================== TMA + MMA issuer (one thread) =====================
    const uint32_t idesc_s = umma_idesc(FMT, AQ, KT);
    const uint32_t idesc_o = umma_idesc(FMT, AQ, nv) | (1u << 16);  // B (= V) is MN-major
    mbar_expect_tx(bar_q, Q_BYTES);
    for (int kb = 0; kb < KB; ++kb)
      tma_load_2d(s_q + kb * (AQ * 128), &map_q, bar_q, p.q_col0 + h * p.hd_stride + kb * 64, b * p.q_len + qb * AQ);
    // request K (and V in pass 1) of flat step i into buffer i & 1
    const int total = 2 * n_tiles;
    auto request = [&](int i) {
      const int pass = i >= n_tiles ? 1 : 0, t = pass ? i - n_tiles : i, bf = i & 1;
      const int row = p.page_table ? p.page_table[(size_t)b * p.pt_stride + t] * KT
                                   : (int)(b * p.kv_batch_rows) + t * KT;
      mbar_expect_tx(bar_kv(bf), pass == 1 ? K_BYTES + V_BYTES : K_BYTES);
      for (int kb = 0; kb < KB; ++kb)
        tma_load_2d(s_k(bf) + kb * (KT * 128), &map_k, bar_kv(bf), p.k_col0 + kvh * p.hd_stride + kb * 64, row);
      if (pass == 1)
        for (int kb = 0; kb < KB; ++kb)
          tma_load_2d(s_v(bf) + kb * (KT * 128), &map_v, bar_kv(bf), p.v_col0 + kvh * p.hd_stride + kb * 64, row);
    };
    if (total > 0) request(0);
    mbar_wait(bar_q, 0);
    for (int it = 0; it < total; ++it) {
      const int pass = it >= n_tiles ? 1 : 0, t = pass ? it - n_tiles : it, bf = it & 1;
      if (it > 0) mbar_wait(bar_sfree, (it - 1) & 1);            // S of the previous tile has been read
      if (pass == 1 && t > 0) mbar_wait(bar_o, (t - 1) & 1);     // previous PV finished with V and P
      mbar_wait(bar_kv(bf), (it >> 1) & 1);
      tc_fence_after();
      for (int ks = 0; ks < k_steps; ++ks) {                     // S = Q K^T
        const uint32_t off_q = (ks / 4) * (AQ * 128) + (ks % 4) * 32, off_k = (ks / 4) * (KT * 128) + (ks % 4) * 32;
        umma(tmem_base + S_COL, umma_desc(s_q + off_q), umma_desc(s_k(bf) + off_k), idesc_s, ks > 0 ? 1u : 0u);
      }
      umma_commit(bar_s);
      // the other buffer held step it-1: its K was consumed by S(it-1) (complete: bar_sfree above), its V by PV(it-1)
      // (complete: bar_o above) -- free for step it+1
      if (it + 1 < total) request(it + 1);
      if (pass == 1) {
        mbar_wait(bar_p, t & 1);                                 // P tile is in shared memory
        tc_fence_after();
        for (int ks = 0; ks < KT / 16; ++ks) {                   // O += P V
          const uint32_t off_p = (ks / 4) * (AQ * 128) + (ks % 4) * 32;
          umma(tmem_base + O_COL, umma_desc(s_p + off_p), umma_desc_mn(s_v(bf) + ks * 2048, KT * 128), idesc_o,
               (t > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(bar_o);
      }
    }
  } else if (warp >= 4) {
    // ===================== softmax + epilogue: thread <-> query row =====================
    const int qw = warp & 3, r = qw * 32 + lane;
    const uint32_t t_row = tmem_base + ((uint32_t)(qw * 32) << 16);
    float m = -INFINITY, l = 0.f;
    int it = 0;
    for (int pass = 0; pass < 2; ++pass) {
      for (int t = 0; t < n_tiles; ++t, ++it) {
        mbar_wait(bar_s, it & 1);
        tc_fence_after();
        if (pass == 1 && t > 0) mbar_wait(bar_o, (t - 1) & 1);     // P buffer is free again
#pragma unroll 1
        for (int c = 0; c < KT / 32; ++c) {
          float s[32];
          tmem_ld32(t_row + S_COL + c * 32, s);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = rnd<T>(s[j]);
            x = (p.scale_mul != 0.f) ? x * p.scale_mul : rnd<T>(p.scale_mode ? x / p.scale : x * p.scale);
            s[j] = (t * KT + c * 32 + j < T_len) ? x : -INFINITY;
          }
          if (pass == 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j) m = fmaxf(m, s[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) { s[j] = __expf(s[j] - m); l += s[j]; }
            // P[r][c*32 .. c*32+31] as bf16 into the K-major SWIZZLE_128B tile: 64-key blocks of 128 rows x 128 B
            uint8_t* blk = p_ptr + (c / 2) * (AQ * 128) + r * 128;
#pragma unroll
            for (int j0 = 0; j0 < 32; j0 += 8) {
              const int chunk = ((c & 1) * 4 + j0 / 8) ^ (r & 7);
              *reinterpret_cast<uint4*>(blk + chunk * 16) = pack<T>(s + j0);
            }
          }
        }
        tc_fence_before();
        if (pass == 1) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> UMMA
          mbar_arrive(bar_p);
        }
        mbar_arrive(bar_sfree);
      }
    }
    if (n_tiles > 0) {
      mbar_wait(bar_o, (n_tiles - 1) & 1);
      tc_fence_after();
      const int qi = qb * AQ + r;
      const float inv = 1.f / l;
      T* orow = reinterpret_cast<T*>(p.out) + (size_t)(b * p.q_len + qi) * p.ld_out + (size_t)h * p.hd;
#pragma unroll 1
      for (int c = 0; c * 32 < p.hd; ++c) {
        float o[32];
        tmem_ld32(t_row + O_COL + c * 32, o);
        tmem_ld_wait();
        if (qi < p.q_len) {
#pragma unroll
          for (int j0 = 0; j0 < 32; j0 += 8) {
            if (c * 32 + j0 >= p.hd) break;  // hd is a multiple of 8
            float w[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) w[j] = o[j0 + j] * inv;
            *reinterpret_cast<uint4*>(orow + c * 32 + j0) = pack<T>(w);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_ALLOC) : "memory");
  }
}


// ------------------------------------------------------------------------------------------------
// ViT attention (SigLIP: <= 256 keys, head_dim <= 80, q/k/v in one fused row-major tensor).
// One CTA = 128 queries of one (image, head); 2 CTAs per SM overlap each other's load / MMA / softmax phases.
//   * only the real head columns are staged: columns 0..63 as a SWIZZLE_128B block, columns 64..79 as a
//     16-column SWIZZLE_32B block (one K step of QK^T, one N=16 slice of PV): 100 KB instead of 160 KB;
//   * S = Q K^T for all keys is ONE 128 x 256 fp32 accumulator in TMEM (5 MMAs); no second QK^T pass;
//   * 8 softmax warps: two threads per query row, each owns 128 key columns; row max of the raw accumulators
//     first (rounding and the positive scale are monotonic), halves exchanged through shared memory, then
//     exp -> bf16 P written over the dead Q/K tiles; O = P V accumulates into the dead S columns.
// 256 TMEM columns and ~106 KB of shared memory per CTA.
constexpr int VIT_TK = 256;
constexpr int VIT_THREADS = 384;

__device__ __forceinline__ uint64_t umma_desc_sw32(uint32_t smem_addr) {
  // K-major, SWIZZLE_32B: rows of 32 B (one 16-element K step), 8-row groups 256 B apart (SBO)
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (16ull << 32) | (1ull << 46) | (6ull << 61);
}
__device__ __forceinline__ uint64_t umma_desc_mn_sw32(uint32_t smem_addr, uint32_t lbo_bytes) {
  // MN-major, SWIZZLE_32B: 16 contiguous MN elements (32 B) per K row, 8-row groups 256 B apart (SBO)
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | (16ull << 32) |
         (1ull << 46) | (6ull << 61);
}
__device__ __forceinline__ void vit_softmax_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
constexpr float LOG2E_F = 1.4426950408889634f;
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <typename T>
__global__ void __launch_bounds__(VIT_THREADS, 2)
attention_vit_kernel(const __grid_constant__ CUtensorMap map64, const __grid_constant__ CUtensorMap map16, AttnParams p) {
  constexpr int FMT = std::is_same<T, bf16>::value ? 1 : 0;
  // shared memory (offsets from a 1024-byte aligned base); P (64 KB) overlays Q and K once S is complete
  constexpr uint32_t OFF_Q0 = 0, OFF_Q1 = 16384, OFF_K0 = 20480, OFF_K1 = 53248, OFF_P = 0;
  constexpr uint32_t OFF_V0 = 65536, OFF_V1 = 98304, OFF_RED = 106496, OFF_BAR = OFF_RED + 2048;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t s_q0 = base + OFF_Q0, s_q1 = base + OFF_Q1, s_k0 = base + OFF_K0, s_k1 = base + OFF_K1;
  const uint32_t s_v0 = base + OFF_V0, s_v1 = base + OFF_V1, s_p = base + OFF_P;
  const uint32_t bar_qk = base + OFF_BAR, bar_v = bar_qk + 8, bar_s = bar_qk + 16, bar_p = bar_qk + 24, bar_o = bar_qk + 32,
                 tmem_slot = bar_qk + 40;
  float* red_m = reinterpret_cast<float*>(base_ptr + OFF_RED);   // [2][128]
  float* red_l = red_m + 256;                                    // [2][128]
  uint8_t* p_ptr = base_ptr + OFF_P;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(base_ptr + OFF_BAR + 40);

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z, kvh = h / p.kv_group;
  const int T_len = p.kv_len_const;
  const bool has_b1 = p.hd > 64;                               // head columns 64..79 exist
  const int k_steps0 = has_b1 ? 4 : (p.hd + 15) / 16;          // 16-wide K steps inside the 64-column block
  const int n_s = ((T_len + 15) / 16) * 16;                    // S columns computed
  const int key_steps = (T_len + 15) / 16;                     // 16-key steps of P V
  const int n_o0 = has_b1 ? 64 : ((p.hd + 15) / 16) * 16;      // PV output columns from the 64-column V block

  if (warp == 1 && lane == 0) {
    mbar_init(bar_qk, 1); mbar_init(bar_v, 1); mbar_init(bar_s, 1); mbar_init(bar_p, 256); mbar_init(bar_o, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0 && lane == 0) {
    // ===================== TMA + MMA issuer (one thread) =====================
    const int row_q = b * p.q_len + qb * AQ, row_kv = (int)(b * p.kv_batch_rows);
    const int cq = p.q_col0 + h * p.hd_stride, ck = p.k_col0 + kvh * p.hd_stride, cv = p.v_col0 + kvh * p.hd_stride;
    pdl_wait();  // q / k / v come from the preceding GEMM
    mbar_expect_tx(bar_qk, has_b1 ? 61440u : 49152u);
    tma_load_2d(s_q0, &map64, bar_qk, cq, row_q);
    tma_load_2d(s_k0, &map64, bar_qk, ck, row_kv);
    tma_load_2d(s_k0 + 16384, &map64, bar_qk, ck, row_kv + 128);
    if (has_b1) {
      tma_load_2d(s_q1, &map16, bar_qk, cq + 64, row_q);
      tma_load_2d(s_k1, &map16, bar_qk, ck + 64, row_kv);
      tma_load_2d(s_k1 + 4096, &map16, bar_qk, ck + 64, row_kv + 128);
    }
    mbar_expect_tx(bar_v, has_b1 ? 40960u : 32768u);
    tma_load_2d(s_v0, &map64, bar_v, cv, row_kv);
    tma_load_2d(s_v0 + 16384, &map64, bar_v, cv, row_kv + 128);
    if (has_b1) {
      tma_load_2d(s_v1, &map16, bar_v, cv + 64, row_kv);
      tma_load_2d(s_v1 + 4096, &map16, bar_v, cv + 64, row_kv + 128);
    }
    mbar_wait(bar_qk, 0);
    tc_fence_after();
    const uint32_t idesc_s = umma_idesc(FMT, AQ, n_s);
    for (int ks = 0; ks < k_steps0; ++ks)                      // S = Q K^T
      umma(tmem_base, umma_desc(s_q0 + ks * 32), umma_desc(s_k0 + ks * 32), idesc_s, ks > 0 ? 1u : 0u);
    if (has_b1) umma(tmem_base, umma_desc_sw32(s_q1), umma_desc_sw32(s_k1), idesc_s, 1u);
    umma_commit(bar_s);
    mbar_wait(bar_v, 0);
    mbar_wait(bar_p, 0);                                       // P complete, every S column has been read
    tc_fence_after();
    const uint32_t idesc_o0 = umma_idesc(FMT, AQ, n_o0) | (1u << 16);  // B (= V) is MN-major
    const uint32_t idesc_o1 = umma_idesc(FMT, AQ, 16) | (1u << 16);
    for (int ks = 0; ks < key_steps; ++ks) {                   // O = P V into the dead S columns
      const uint64_t a = umma_desc(s_p + (ks / 4) * (AQ * 128) + (ks % 4) * 32);
      umma(tmem_base, a, umma_desc_mn(s_v0 + ks * 2048, VIT_TK * 128), idesc_o0, ks > 0 ? 1u : 0u);
      if (has_b1) umma(tmem_base + 64, a, umma_desc_mn_sw32(s_v1 + ks * 512, VIT_TK * 32), idesc_o1, ks > 0 ? 1u : 0u);
    }
    umma_commit(bar_o);
  } else if (warp >= 4) {
    // ===================== softmax + epilogue: two threads per query row =====================
    const int q4 = warp & 3, hf = (warp - 4) >> 2, r = q4 * 32 + lane;
    const uint32_t t_row = tmem_base + ((uint32_t)(q4 * 32) << 16);
    const int col0 = hf * 128;
    const float mul = p.scale_mul != 0.f ? p.scale_mul : p.scale;  // exact power of two, or s*scale rounded to T
    float s[32];
    pdl_wait();  // the output rows may still be read by a running predecessor
    mbar_wait(bar_s, 0);
    tc_fence_after();
    float m = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      const int cb = col0 + c * 32;
      if (cb >= T_len) break;
      tmem_ld32(t_row + cb, s);
      tmem_ld_wait();
      if (cb + 32 <= T_len) {
#pragma unroll
        for (int j = 0; j < 32; ++j) m = fmaxf(m, s[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) m = fmaxf(m, cb + j < T_len ? s[j] : -INFINITY);
      }
    }
    red_m[hf * 128 + r] = m;
    vit_softmax_sync();
    m = fmaxf(red_m[r], red_m[128 + r]);
    m = rnd<T>(rnd<T>(m) * mul);                               // the row max after the reference's roundings
    const float neg_m_l2 = -m * LOG2E_F;
    float l = 0.f;
    const int t_pad = ((T_len + 63) / 64) * 64;                // P columns the MMAs may touch
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      const int cb = col0 + c * 32;
      if (cb >= t_pad) break;
      if (cb < T_len) {
        tmem_ld32(t_row + cb, s);
        tmem_ld_wait();
        // p = exp(x - m) as ex2(x * log2e - m * log2e): one FFMA + one MUFU.EX2 per element (results below 2^-126
        // flush to zero); x = QK^T rounded to T, scaled, rounded again -- two values per packed conversion
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          float x0 = s[j], x1 = s[j + 1];
          rnd2<T>(x0, x1);
          x0 *= mul; x1 *= mul;
          rnd2<T>(x0, x1);
          s[j] = ex2_ftz(fmaf(x0, LOG2E_F, neg_m_l2));
          s[j + 1] = ex2_ftz(fmaf(x1, LOG2E_F, neg_m_l2));
        }
        if (cb + 32 > T_len) {                                 // ragged last chunk only
#pragma unroll
          for (int j = 0; j < 32; ++j) s[j] = (cb + j < T_len) ? s[j] : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 32; j += 2) { l0 += s[j]; l1 += s[j + 1]; }
        l += l0 + l1;
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) s[j] = 0.f;
      }
      // P[r][cb .. cb+31] into the K-major SWIZZLE_128B tile: 64-key blocks of 128 rows x 128 B
      uint8_t* blk = p_ptr + (cb / 64) * (AQ * 128) + r * 128;
#pragma unroll
      for (int j0 = 0; j0 < 32; j0 += 8) {
        const int chunk = (((cb / 32) & 1) * 4 + j0 / 8) ^ (r & 7);
        *reinterpret_cast<uint4*>(blk + chunk * 16) = pack<T>(s + j0);
      }
    }
    red_l[hf * 128 + r] = l;
    tc_fence_before();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> UMMA
    mbar_arrive(bar_p);
    vit_softmax_sync();
    const float inv = 1.f / (red_l[r] + red_l[128 + r]);
    mbar_wait(bar_o, 0);
    tc_fence_after();
    const int qi = qb * AQ + r;
    T* orow = reinterpret_cast<T*>(p.out) + (size_t)(b * p.q_len + qi) * p.ld_out + (size_t)h * p.hd;
    // output columns: this thread's half takes [0,48) or [48,80)
    const int oc0 = hf ? 48 : 0;
    if (oc0 < p.hd) {
      tmem_ld32(t_row + oc0, s);
      float s2[16];
      if (!hf) tmem_ld16(t_row + 32, s2);
      tmem_ld_wait();
      if (qi < p.q_len) {
#pragma unroll
        for (int j0 = 0; j0 < 32; j0 += 8) {
          if (oc0 + j0 < p.hd) {                               // hd is a multiple of 8
            float w[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) w[j] = s[j0 + j] * inv;
            *reinterpret_cast<uint4*>(orow + oc0 + j0) = pack<T>(w);
          }
        }
        if (!hf) {
#pragma unroll
          for (int j0 = 0; j0 < 16; j0 += 8) {
            if (32 + j0 < p.hd) {
              float w[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) w[j] = s2[j0 + j] * inv;
              *reinterpret_cast<uint4*>(orow + 32 + j0) = pack<T>(w);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
  }
}

template <typename T>
static int launch_attn_vit(const CUtensorMap& m64, const CUtensorMap& m16, const AttnParams& p, int B, cudaStream_t st) {
  const size_t smem = 1024 + 106496 + 2048 + 64;
  auto kern = attention_vit_kernel<T>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    set_error("attention_tc(vit): cannot reserve %zu B of shared memory", smem);
    cudaGetLastError();
    return PG_ERR_CUDA;
  }
  const int ctas = cdiv(p.q_len, AQ) * p.n_heads * B;
  return launch_tc("attention_vit", kern, dim3(cdiv(p.q_len, AQ), p.n_heads, B), dim3(VIT_THREADS), smem, 1, ctas <= 296, st, m64, m16, p);
}

template <typename T, int HDP, int KT>
static int launch_attn(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const AttnParams& p, int B,
                       cudaStream_t st) {
  const size_t smem = 1024 + (size_t)(AQ * HDP * 2) + 4 * (size_t)(KT * HDP * 2) + (size_t)(AQ * KT * 2) + 64;
  auto kern = attention_tc_kernel<T, HDP, KT>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    set_error("attention_tc: cannot reserve %zu B of shared memory", smem);
    cudaGetLastError();
    return PG_ERR_CUDA;
  }
  const int ctas = cdiv(p.q_len, AQ) * p.n_heads * B;
  return launch_tc("attention_tc", kern, dim3(cdiv(p.q_len, AQ), p.n_heads, B), dim3(256), smem, 1, ctas <= 296, st, mq, mk, mv, p);
}

}  // namespace tc
}  // namespace pg

using namespace pg;

// Tensor-core attention.  q/k/v: 16-bit row-major matrices with `*_rows` rows of `ld_*` elements; head h of
// Q lives at columns q_col0 + h*hd_stride .. +hd (columns hd..hd_stride-1 of a padded head must be zero
// or belong to memory that multiplies to zero: the tiled kernel loads ceil(hd/64)*64 columns,
// the ViT kernel multiplies ceil(hd/16)*16 columns).
// Output: [B*q_len, ld_out] with head h at column h*hd (packed).
extern "C" int pg_attention_tc(void* out, int ld_out, const void* q, long long q_rows, int ld_q, int q_col0,
                               const void* k, const void* v, long long kv_rows, int ld_kv, int k_col0, int v_col0,
                               int hd_stride, long long kv_batch_rows, const int32_t* page_table, int pt_stride,
                               int page_size, const int32_t* kv_len, int kv_len_const, int kv_len_add, int B, int q_len,
                               int n_heads, int n_kv_heads, int hd, float scale, int scale_mode, int dtype,
                               void* stream) {
  PG_REQUIRE(dtype == PG_BF16 || dtype == PG_F16, "attention_tc: 16-bit dtypes only");
  PG_REQUIRE(hd % 8 == 0 && hd <= 256 && hd_stride >= hd && hd_stride % 8 == 0, "attention_tc: unsupported head_dim %d / stride %d", hd, hd_stride);
  PG_REQUIRE(ld_q % 8 == 0 && ld_kv % 8 == 0 && ld_out % 8 == 0 && q_col0 % 8 == 0 && k_col0 % 8 == 0 && v_col0 % 8 == 0,
             "attention_tc: rows must be 16-byte aligned");
  PG_REQUIRE(n_heads % n_kv_heads == 0, "attention_tc: bad head counts");
  const int hdp = hd <= 128 ? 128 : 256;
  const int kt = 64;
  PG_REQUIRE(!page_table || page_size == kt, "attention_tc: page size must equal the key tile (%d)", kt);
  const bool bf = dtype == PG_BF16;
  tc::AttnParams p = {out, ld_out, hd, q_len, n_heads, n_heads / n_kv_heads, q_col0, k_col0, v_col0, hd_stride,
                      kv_batch_rows, page_table, pt_stride, kv_len, kv_len_const, kv_len_add, scale, scale_mode, 0.f};
  int ex = 0;
  if (scale_mode == 1 && frexpf(scale, &ex) == 0.5f) p.scale_mul = 1.0f / scale;  // x / 2^k == x * 2^-k exactly
  cudaStream_t st = (cudaStream_t)stream;
  // SigLIP-shaped problems: all keys at once, real head columns only (see attention_vit_kernel)
  static const int vit = env_int("PG_ATTN_VIT", 1);
  const int hd16 = ((hd + 15) / 16) * 16;
  if (vit && hd <= 80 && (hd == hd16 || hd_stride >= hd16) && !page_table && !kv_len && kv_len_const >= 1 &&
      kv_len_const <= tc::VIT_TK && q == k && k == v && ld_q == ld_kv && q_rows == kv_rows &&
      (scale_mode == 0 || p.scale_mul != 0.f)) {
    CUtensorMap m64, m16;
    PG_REQUIRE(tc::make_map_2d_ex(&m64, q, q_rows, ld_q, ld_q, 64, tc::AQ, 128, bf) &&
                   tc::make_map_2d_ex(&m16, q, q_rows, ld_q, ld_q, 16, tc::AQ, 32, bf),
               "attention_tc(vit): cuTensorMapEncodeTiled failed");
    return bf ? tc::launch_attn_vit<bf16>(m64, m16, p, B, st) : tc::launch_attn_vit<f16>(m64, m16, p, B, st);
  }
  CUtensorMap mq, mk, mv;
  PG_REQUIRE(tc::make_map_2d(&mq, q, q_rows, ld_q, ld_q, tc::AQ, bf) && tc::make_map_2d(&mk, k, kv_rows, ld_kv, ld_kv, kt, bf) &&
                 tc::make_map_2d(&mv, v, kv_rows, ld_kv, ld_kv, kt, bf),
             "attention_tc: cuTensorMapEncodeTiled failed");
  PG_REQUIRE(hd == hdp || hd_stride >= ((hd + 63) / 64) * 64 || n_heads == 1, "attention_tc: head rows must be padded to a multiple of 64 columns");
  if (hdp == 128) return bf ? tc::launch_attn<bf16, 128, 64>(mq, mk, mv, p, B, st) : tc::launch_attn<f16, 128, 64>(mq, mk, mv, p, B, st);
  return bf ? tc::launch_attn<bf16, 256, 64>(mq, mk, mv, p, B, st) : tc::launch_attn<f16, 256, 64>(mq, mk, mv, p, B, st);
}
