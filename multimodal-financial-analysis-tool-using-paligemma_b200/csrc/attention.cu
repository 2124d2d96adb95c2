// Attention kernels.  The reference never masks (additive mask is all zeros,
// modeling_gemma.py:506-514; SigLIP has no mask), softmax runs in fp32 (:273, siglip :125).
//   (decode, q_len == 1: decode_attention.cu)
//   attention_kernel        : general q_len x kv_len (prefill, cache-off recompute, SigLIP), K/V
//                             contiguous or paged, flash-style online softmax, SIMT fp32.
#include "common.cuh"

namespace pg {

// ------------------------------------------------------------------------------------------
// General attention: CTA = AT_WARPS query rows of one (b, head); each warp owns one query.
// Keys are visited in tiles of 32 staged in shared memory as fp32: lane j scores key j, then the
// probabilities are broadcast by shuffle for the PV accumulation (lane owns dims lane+32c).
constexpr int AT_WARPS = 8;
constexpr int AT_TILE = 32;

template <typename T, int NC>  // NC = ceil(hd/32)
__global__ void __launch_bounds__(AT_WARPS * 32)
attention_kernel(T* __restrict__ out, int ld_out, const T* __restrict__ q, int ld_q, const T* __restrict__ k,
                 const T* __restrict__ v, int ld_kv, long long kv_batch_stride,
                 const int32_t* __restrict__ page_table, int pt_stride, int page_size,
                 const int32_t* __restrict__ kv_len_dev, int kv_len_const, int kv_len_add, int q_len,
                 int n_heads, int n_kv_heads, int hd, float scale, int scale_mode) {
  constexpr int V = Vec<T>::N;
  const int b = blockIdx.z, h = blockIdx.y, kvh = h / (n_heads / n_kv_heads);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int qi = blockIdx.x * AT_WARPS + wid;
  const int T_len = kv_len_dev ? (kv_len_dev[b] + kv_len_add) : kv_len_const;
  const int hdp = hd + 4;  // padded K row: float4 reads by 32 lanes hit distinct banks per phase

  extern __shared__ __align__(16) float sm[];
  float* s_k = sm;                                  // [32][hd+4]
  float* s_v = s_k + AT_TILE * hdp;                 // [32][hd]
  float* s_q = s_v + AT_TILE * hd;                  // [AT_WARPS][hd]

  if (qi < q_len) {
    const T* qrow = q + (size_t)(b * q_len + qi) * ld_q + (size_t)h * hd;
    for (int d = lane; d < hd; d += 32) s_q[wid * hd + d] = to_f<T>(qrow[d]);
  }
  float m = -INFINITY, l = 0.f, acc[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) acc[c] = 0.f;

  const int vec_per_row = hd / V;
  for (int j0 = 0; j0 < T_len; j0 += AT_TILE) {
    __syncthreads();  // previous tile fully consumed (also orders the s_q writes)
    for (int e = threadIdx.x; e < AT_TILE * vec_per_row; e += AT_WARPS * 32) {
      const int r = e / vec_per_row, d = (e % vec_per_row) * V;
      const int j = j0 + r;
      float kf[V], vf[V];
      if (j < T_len) {
        size_t row;
        if (page_table) {
          const int page = page_table[(size_t)b * pt_stride + j / page_size];
          row = ((size_t)page * page_size + (j % page_size)) * (size_t)(n_kv_heads * hd) + (size_t)kvh * hd;
        } else {
          row = (size_t)b * kv_batch_stride + (size_t)j * ld_kv + (size_t)kvh * hd;
        }
        unpack<T>(ldg_cached(k + row + d), kf);
        unpack<T>(ldg_cached(v + row + d), vf);
      } else {
#pragma unroll
        for (int i = 0; i < V; ++i) { kf[i] = 0.f; vf[i] = 0.f; }
      }
#pragma unroll
      for (int i = 0; i < V; ++i) { s_k[r * hdp + d + i] = kf[i]; s_v[r * hd + d + i] = vf[i]; }
    }
    __syncthreads();
    if (qi < q_len) {
      float s = 0.f;
      const float* kr = s_k + lane * hdp;
      const float* qr = s_q + wid * hd;
      for (int d = 0; d < hd; d += 4) {
        const float4 a = *reinterpret_cast<const float4*>(qr + d);
        const float4 kk = *reinterpret_cast<const float4*>(kr + d);
        s = fmaf(a.x, kk.x, s); s = fmaf(a.y, kk.y, s); s = fmaf(a.z, kk.z, s); s = fmaf(a.w, kk.w, s);
      }
      s = rnd<T>(s);
      s = rnd<T>(scale_mode ? s / scale : s * scale);
      if (j0 + lane >= T_len) s = -INFINITY;
      const float mn = fmaxf(m, warp_max(s));
      const float corr = __expf(m - mn);
      const float p = __expf(s - mn);
      l = l * corr + warp_sum(p);
      m = mn;
#pragma unroll
      for (int c = 0; c < NC; ++c) acc[c] *= corr;
      const int jmax = min(AT_TILE, T_len - j0);
      for (int j = 0; j < jmax; ++j) {
        const float pj = __shfl_sync(0xffffffffu, p, j);
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const int d = lane + 32 * c;
          if (d < hd) acc[c] = fmaf(pj, s_v[j * hd + d], acc[c]);
        }
      }
    }
  }
  if (qi < q_len) {
    T* orow = out + (size_t)(b * q_len + qi) * ld_out + (size_t)h * hd;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int d = lane + 32 * c;
      if (d < hd) orow[d] = from_f<T>(acc[c] / l);
    }
  }
}

}  // namespace pg

using namespace pg;

extern "C" {

int pg_attention(void* out, int ld_out, const void* q, int ld_q, const void* k, const void* v, int ld_kv,
                 long long kv_batch_stride, const int32_t* page_table, int pt_stride, int page_size,
                 const int32_t* kv_len, int kv_len_const, int kv_len_add, int B, int q_len, int n_heads,
                 int n_kv_heads, int hd, float scale, int scale_mode, int dtype, void* stream) {
  if (B * q_len <= 0) return PG_OK;
  PG_REQUIRE(n_heads % n_kv_heads == 0, "attention: heads %d not a multiple of kv heads %d", n_heads, n_kv_heads);
  dim3 grid(cdiv(q_len, AT_WARPS), n_heads, B);
  const size_t smem = ((size_t)AT_TILE * (hd + 4) + (size_t)AT_TILE * hd + (size_t)AT_WARPS * hd) * sizeof(float);
#define PG_AT_LAUNCH(NC)                                                                              \
  {                                                                                                   \
    auto kern = attention_kernel<T, NC>;                                                              \
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    kern<<<grid, AT_WARPS * 32, smem, (cudaStream_t)stream>>>(                                        \
        (T*)out, ld_out, (const T*)q, ld_q, (const T*)k, (const T*)v, ld_kv, kv_batch_stride,         \
        page_table, pt_stride, page_size, kv_len, kv_len_const, kv_len_add, q_len, n_heads,           \
        n_kv_heads, hd, scale, scale_mode);                                                           \
  }
  PG_DISPATCH_DTYPE(dtype, T, {
    constexpr int V = Vec<T>::N;
    PG_REQUIRE(hd % V == 0 && hd % 4 == 0 && hd <= 256, "attention: unsupported head_dim %d", hd);
    PG_REQUIRE(ld_kv % V == 0 && kv_batch_stride % V == 0, "attention: K/V rows not 16-byte aligned");
    const int nc = (hd + 31) / 32;
    switch (nc) {
      case 1: PG_AT_LAUNCH(1) break;
      case 2: PG_AT_LAUNCH(2) break;
      case 3: PG_AT_LAUNCH(3) break;
      case 4: PG_AT_LAUNCH(4) break;
      case 5: PG_AT_LAUNCH(5) break;
      case 6: PG_AT_LAUNCH(6) break;
      case 7: PG_AT_LAUNCH(7) break;
      default: PG_AT_LAUNCH(8) break;
    }
  });
#undef PG_AT_LAUNCH
  return check_launch("attention");
}

}  // extern "C"
