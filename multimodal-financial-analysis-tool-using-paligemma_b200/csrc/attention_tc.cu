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
  const uint32_t s_q = base, s_k = s_q + Q_BYTES, s_v = s_k + K_BYTES, s_p = s_v + V_BYTES;
  const uint32_t bars = s_p + P_BYTES;
  const uint32_t bar_q = bars, bar_kv = bars + 8, bar_s = bars + 16, bar_sfree = bars + 24, bar_p = bars + 32,
                 bar_o = bars + 40, tmem_slot = bars + 48;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  uint8_t* p_ptr = smem_raw + (s_p - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z, kvh = h / p.kv_group;
  const int T_len = p.kv_len_dev ? (p.kv_len_dev[b] + p.kv_len_add) : p.kv_len_const;
  const int n_tiles = (T_len + KT - 1) / KT;
  const int nv = ((p.hd + 15) / 16) * 16;        // PV output columns (multiple of 16)
  const int k_steps = (p.hd + 15) / 16;          // 16-wide K steps of QK^T that hold real data

  if (warp == 1 && lane == 0) {
    mbar_init(bar_q, 1); mbar_init(bar_kv, 1); mbar_init(bar_s, 1); mbar_init(bar_sfree, 128);
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

  if (warp == 0 && lane == 0) {
    // ===================== TMA + MMA issuer (one thread) =====================
    const uint32_t idesc_s = umma_idesc(FMT, AQ, KT);
    const uint32_t idesc_o = umma_idesc(FMT, AQ, nv) | (1u << 16);  // B (= V) is MN-major
    mbar_expect_tx(bar_q, Q_BYTES);
    for (int kb = 0; kb < KB; ++kb)
      tma_load_2d(s_q + kb * (AQ * 128), &map_q, bar_q, p.q_col0 + h * p.hd_stride + kb * 64, b * p.q_len + qb * AQ);
    mbar_wait(bar_q, 0);
    int it = 0;
    for (int pass = 0; pass < 2; ++pass) {
      for (int t = 0; t < n_tiles; ++t, ++it) {
        if (it > 0) mbar_wait(bar_sfree, (it - 1) & 1);            // S of the previous tile has been read
        if (pass == 1 && t > 0) mbar_wait(bar_o, (t - 1) & 1);     // previous PV finished with V and P
        const int row = p.page_table ? p.page_table[(size_t)b * p.pt_stride + t] * KT
                                     : (int)(b * p.kv_batch_rows) + t * KT;
        mbar_expect_tx(bar_kv, pass == 1 ? K_BYTES + V_BYTES : K_BYTES);
        for (int kb = 0; kb < KB; ++kb)
          tma_load_2d(s_k + kb * (KT * 128), &map_k, bar_kv, p.k_col0 + kvh * p.hd_stride + kb * 64, row);
        if (pass == 1)
          for (int kb = 0; kb < KB; ++kb)
            tma_load_2d(s_v + kb * (KT * 128), &map_v, bar_kv, p.v_col0 + kvh * p.hd_stride + kb * 64, row);
        mbar_wait(bar_kv, it & 1);
        tc_fence_after();
        for (int ks = 0; ks < k_steps; ++ks) {                     // S = Q K^T
          const uint32_t off_q = (ks / 4) * (AQ * 128) + (ks % 4) * 32, off_k = (ks / 4) * (KT * 128) + (ks % 4) * 32;
          umma(tmem_base + S_COL, umma_desc(s_q + off_q), umma_desc(s_k + off_k), idesc_s, ks > 0 ? 1u : 0u);
        }
        umma_commit(bar_s);
        if (pass == 1) {
          mbar_wait(bar_p, t & 1);                                 // P tile is in shared memory
          tc_fence_after();
          for (int ks = 0; ks < KT / 16; ++ks) {                   // O += P V
            const uint32_t off_p = (ks / 4) * (AQ * 128) + (ks % 4) * 32;
            umma(tmem_base + O_COL, umma_desc(s_p + off_p), umma_desc_mn(s_v + ks * 2048, KT * 128), idesc_o,
                 (t > 0 || ks > 0) ? 1u : 0u);
          }
          umma_commit(bar_o);
        }
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
// One-shot variant for short, contiguous key sets (SigLIP: 256 patches, head_dim <= 128): all keys and
// values of the (batch, head) are staged at once, S = Q K^T is a single 128 x 256 accumulator that stays
// in TMEM (read twice: row max, then exp), P overwrites the K tile in shared memory, O = P V is one
// more chain of MMAs.  Three dependent steps per CTA instead of a per-tile loop.
template <typename T>
__global__ void __launch_bounds__(256, 1)
attention_tc_oneshot_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv,
                            AttnParams p) {
  constexpr int HDP = 128, KB = 2, TK = 256;                 // padded head dim, 64-column blocks, keys staged
  constexpr int Q_BYTES = AQ * HDP * 2, KV_BYTES = TK * HDP * 2;  // 32 KB, 64 KB
  constexpr int S_COL = 0, O_COL = TK;
  constexpr int FMT = std::is_same<T, bf16>::value ? 1 : 0;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t s_q = base, s_k = s_q + Q_BYTES, s_v = s_k + KV_BYTES, s_p = s_k;  // P reuses the K tile
  const uint32_t bars = s_v + KV_BYTES;
  const uint32_t bar_ld = bars, bar_s = bars + 8, bar_p = bars + 16, bar_o = bars + 24, tmem_slot = bars + 32;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  uint8_t* p_ptr = smem_raw + (s_p - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z, kvh = h / p.kv_group;
  const int T_len = p.kv_len_const;
  const int nv = ((p.hd + 15) / 16) * 16, k_steps = (p.hd + 15) / 16;
  const int key_steps = (T_len + 15) / 16;                   // 16-key steps of P V that hold real keys

  if (warp == 1 && lane == 0) {
    mbar_init(bar_ld, 1); mbar_init(bar_s, 1); mbar_init(bar_p, 128); mbar_init(bar_o, 1);
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
    const uint32_t idesc_s = umma_idesc(FMT, AQ, TK);
    const uint32_t idesc_o = umma_idesc(FMT, AQ, nv) | (1u << 16);
    const int row = (int)(b * p.kv_batch_rows);
    mbar_expect_tx(bar_ld, Q_BYTES + 2 * KV_BYTES);
    for (int kb = 0; kb < KB; ++kb) {
      tma_load_2d(s_q + kb * (AQ * 128), &map_q, bar_ld, p.q_col0 + h * p.hd_stride + kb * 64, b * p.q_len + qb * AQ);
      tma_load_2d(s_k + kb * (TK * 128), &map_kv, bar_ld, p.k_col0 + kvh * p.hd_stride + kb * 64, row);
      tma_load_2d(s_v + kb * (TK * 128), &map_kv, bar_ld, p.v_col0 + kvh * p.hd_stride + kb * 64, row);
    }
    mbar_wait(bar_ld, 0);
    tc_fence_after();
    for (int ks = 0; ks < k_steps; ++ks) {
      const uint32_t off_q = (ks / 4) * (AQ * 128) + (ks % 4) * 32, off_k = (ks / 4) * (TK * 128) + (ks % 4) * 32;
      umma(tmem_base + S_COL, umma_desc(s_q + off_q), umma_desc(s_k + off_k), idesc_s, ks > 0 ? 1u : 0u);
    }
    umma_commit(bar_s);
    mbar_wait(bar_p, 0);
    tc_fence_after();
    for (int ks = 0; ks < key_steps; ++ks) {
      const uint32_t off_p = (ks / 4) * (AQ * 128) + (ks % 4) * 32;
      umma(tmem_base + O_COL, umma_desc(s_p + off_p), umma_desc_mn(s_v + ks * 2048, TK * 128), idesc_o, ks > 0 ? 1u : 0u);
    }
    umma_commit(bar_o);
  } else if (warp >= 4) {
    const int qw = warp & 3, r = qw * 32 + lane;
    const uint32_t t_row = tmem_base + ((uint32_t)(qw * 32) << 16);
    const int n_chunks = (T_len + 31) / 32;
    float m = -INFINITY, l = 0.f;
    mbar_wait(bar_s, 0);
    tc_fence_after();
    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll 1
      for (int c = 0; c < n_chunks; ++c) {
        float s[32];
        tmem_ld32(t_row + S_COL + c * 32, s);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = rnd<T>(s[j]);
          x = (p.scale_mul != 0.f) ? x * p.scale_mul : rnd<T>(p.scale_mode ? x / p.scale : x * p.scale);
          s[j] = (c * 32 + j < T_len) ? x : -INFINITY;
        }
        if (pass == 0) {
#pragma unroll
          for (int j = 0; j < 32; ++j) m = fmaxf(m, s[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) { s[j] = __expf(s[j] - m); l += s[j]; }
          uint8_t* blk = p_ptr + (c / 2) * (AQ * 128) + r * 128;
#pragma unroll
          for (int j0 = 0; j0 < 32; j0 += 8) {
            const int chunk = ((c & 1) * 4 + j0 / 8) ^ (r & 7);
            *reinterpret_cast<uint4*>(blk + chunk * 16) = pack<T>(s + j0);
          }
        }
      }
      if (pass == 1 && (n_chunks & 1)) {
        // the last 16-key MMA step may read the other half of a 64-key block row: keep it finite (zero)
        uint8_t* blk = p_ptr + (n_chunks / 2) * (AQ * 128) + r * 128;
#pragma unroll
        for (int j0 = 0; j0 < 32; j0 += 8)
          *reinterpret_cast<uint4*>(blk + ((4 + j0 / 8) ^ (r & 7)) * 16) = make_uint4(0, 0, 0, 0);
      }
    }
    tc_fence_before();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_arrive(bar_p);
    mbar_wait(bar_o, 0);
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
          if (c * 32 + j0 >= p.hd) break;
          float w[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) w[j] = o[j0 + j] * inv;
          *reinterpret_cast<uint4*>(orow + c * 32 + j0) = pack<T>(w);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

template <typename T>
static int launch_attn_oneshot(const CUtensorMap& mq, const CUtensorMap& mkv, const AttnParams& p, int B, cudaStream_t st) {
  const size_t smem = 1024 + 32768 + 2 * 65536 + 64;
  auto kern = attention_tc_oneshot_kernel<T>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    set_error("attention_tc(oneshot): cannot reserve %zu B of shared memory", smem);
    cudaGetLastError();
    return PG_ERR_CUDA;
  }
  dim3 grid(cdiv(p.q_len, AQ), p.n_heads, B);
  kern<<<grid, 256, smem, st>>>(mq, mkv, p);
  return check_launch("attention_tc_oneshot");
}

template <typename T, int HDP, int KT>
static int launch_attn(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const AttnParams& p, int B,
                       cudaStream_t st) {
  const size_t smem = 1024 + (size_t)(AQ * HDP * 2) + 2 * (size_t)(KT * HDP * 2) + (size_t)(AQ * KT * 2) + 64;
  auto kern = attention_tc_kernel<T, HDP, KT>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    set_error("attention_tc: cannot reserve %zu B of shared memory", smem);
    cudaGetLastError();
    return PG_ERR_CUDA;
  }
  dim3 grid(cdiv(p.q_len, AQ), p.n_heads, B);
  kern<<<grid, 256, smem, st>>>(mq, mk, mv, p);
  return check_launch("attention_tc");
}

}  // namespace tc
}  // namespace pg

using namespace pg;

// Tensor-core attention.  q/k/v: 16-bit row-major matrices with `*_rows` rows of `ld_*` elements; head h of
// Q lives at columns q_col0 + h*hd_stride .. +hd (columns hd..hd_stride-1 of a padded head must be zero
// or belong to memory that multiplies to zero: the kernel always loads ceil(hd/64)*64 columns).
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
  PG_REQUIRE(hd == hdp || hd_stride >= ((hd + 63) / 64) * 64 || n_heads == 1, "attention_tc: head rows must be padded to a multiple of 64 columns");
  PG_REQUIRE(!page_table || page_size == kt, "attention_tc: page size must equal the key tile (%d)", kt);
  const bool bf = dtype == PG_BF16;
  CUtensorMap mq, mk, mv;
  PG_REQUIRE(tc::make_map_2d(&mq, q, q_rows, ld_q, ld_q, tc::AQ, bf) && tc::make_map_2d(&mk, k, kv_rows, ld_kv, ld_kv, kt, bf) &&
                 tc::make_map_2d(&mv, v, kv_rows, ld_kv, ld_kv, kt, bf),
             "attention_tc: cuTensorMapEncodeTiled failed");
  tc::AttnParams p = {out, ld_out, hd, q_len, n_heads, n_heads / n_kv_heads, q_col0, k_col0, v_col0, hd_stride,
                      kv_batch_rows, page_table, pt_stride, kv_len, kv_len_const, kv_len_add, scale, scale_mode, 0.f};
  int ex = 0;
  if (scale_mode == 1 && frexpf(scale, &ex) == 0.5f) p.scale_mul = 1.0f / scale;  // x / 2^k == x * 2^-k exactly
  cudaStream_t st = (cudaStream_t)stream;
  static const int oneshot = env_int("PG_ATTN_ONESHOT", 0);  // measured slower than the tiled kernel at 2 CTAs/SM (29.5 vs 34.4 ms, batch 64)
  if (oneshot && hdp == 128 && !page_table && !kv_len && kv_len_const <= 256 && k == v) {
    CUtensorMap mkv;
    PG_REQUIRE(tc::make_map_2d(&mkv, k, kv_rows, ld_kv, ld_kv, 256, bf), "attention_tc: cuTensorMapEncodeTiled failed");
    return bf ? tc::launch_attn_oneshot<bf16>(mq, mkv, p, B, st) : tc::launch_attn_oneshot<f16>(mq, mkv, p, B, st);
  }
  if (hdp == 128) return bf ? tc::launch_attn<bf16, 128, 64>(mq, mk, mv, p, B, st) : tc::launch_attn<f16, 128, 64>(mq, mk, mv, p, B, st);
  return bf ? tc::launch_attn<bf16, 256, 64>(mq, mk, mv, p, B, st) : tc::launch_attn<f16, 256, 64>(mq, mk, mv, p, B, st);
}
