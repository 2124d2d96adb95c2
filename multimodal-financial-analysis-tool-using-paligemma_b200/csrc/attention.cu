// Attention kernels.  The reference never masks (additive mask is all zeros,
// modeling_gemma.py:506-514; SigLIP has no mask), softmax runs in fp32 (:273, siglip :125).
//   decode_attention_kernel : one new token per sequence, split-K over the paged KV cache, the G
//                             query heads of a KV head share every K/V load (MQA, :136-141,262-263
//                             never materialised); partials merged by the last CTA to finish.
//   attention_kernel        : general q_len x kv_len (prefill, cache-off recompute, SigLIP), K/V
//                             contiguous or paged, flash-style online softmax, SIMT fp32.
#include "common.cuh"

namespace pg {

// ------------------------------------------------------------------------------------------
constexpr int DA_WARPS = 4;
constexpr int DA_MIN_CHUNK = 32;  // tokens per split at least

__device__ __forceinline__ int da_active_splits(int T, int max_splits) {
  int n = (T + DA_MIN_CHUNK - 1) / DA_MIN_CHUNK;
  return n < 1 ? 1 : (n > max_splits ? max_splits : n);
}

// ws layout per (b, kvh, split): G*hd fp32 accumulators, then G (m) and G (l).
template <typename T, int G, int NCH>
__global__ void __launch_bounds__(DA_WARPS * 32)
decode_attention_kernel(T* __restrict__ out, const T* __restrict__ q, const T* __restrict__ k_pool,
                        const T* __restrict__ v_pool, const int32_t* __restrict__ page_table, int pt_stride,
                        int page_size, const int32_t* __restrict__ kv_len, int kv_len_add, int nq, int nkv,
                        int hd, float scale_div, float* __restrict__ ws, int* __restrict__ counters,
                        int max_splits) {
  constexpr int V = Vec<T>::N;
  const int split = blockIdx.x, b = blockIdx.y, kvh = blockIdx.z;
  const int T_len = kv_len[b] + kv_len_add;
  const int n_active = da_active_splits(T_len, max_splits);
  if (split >= n_active) return;
  const int chunk = (T_len + n_active - 1) / n_active;
  const int t0 = split * chunk, t1 = min(T_len, t0 + chunk);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int row_elems = nkv * hd;

  extern __shared__ __align__(16) float sm[];  // [DA_WARPS][G][hd] acc, then [DA_WARPS][G] m, l
  float* s_acc = sm;
  float* s_m = sm + (size_t)DA_WARPS * G * hd;
  float* s_l = s_m + DA_WARPS * G;
  __shared__ int s_ticket;

  // this lane's slice of the G query vectors
  float qf[G][NCH][V];
#pragma unroll
  for (int g = 0; g < G; ++g)
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int d = (c * 32 + lane) * V;
      if (d < hd) unpack<T>(ldg_cached(q + (size_t)b * nq * hd + (size_t)(kvh * G + g) * hd + d), qf[g][c]);
      else
#pragma unroll
        for (int i = 0; i < V; ++i) qf[g][c][i] = 0.f;
    }
  float m[G], l[G], acc[G][NCH][V];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    m[g] = -INFINITY; l[g] = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
      for (int i = 0; i < V; ++i) acc[g][c][i] = 0.f;
  }

  for (int j = t0 + wid; j < t1; j += DA_WARPS) {
    const int page = page_table[(size_t)b * pt_stride + j / page_size];
    const size_t row = ((size_t)page * page_size + (j % page_size)) * (size_t)row_elems + (size_t)kvh * hd;
    float kf[NCH][V], vf[NCH][V];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int d = (c * 32 + lane) * V;
      if (d < hd) {
        unpack<T>(ldg_stream(k_pool + row + d), kf[c]);
        unpack<T>(ldg_stream(v_pool + row + d), vf[c]);
      } else {
#pragma unroll
        for (int i = 0; i < V; ++i) { kf[c][i] = 0.f; vf[c][i] = 0.f; }
      }
    }
    float s[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      float t = 0.f;
#pragma unroll
      for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int i = 0; i < V; ++i) t = fmaf(qf[g][c][i], kf[c][i], t);
      s[g] = t;
    }
#pragma unroll
    for (int g = 0; g < G; ++g) s[g] = warp_sum(s[g]);
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const float sc = rnd<T>(rnd<T>(s[g]) / scale_div);  // matmul output, then "/ sqrt(hd)" (:266)
      const float mn = fmaxf(m[g], sc);
      const float corr = __expf(m[g] - mn);
      const float p = __expf(sc - mn);
      l[g] = l[g] * corr + p;
      m[g] = mn;
#pragma unroll
      for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int i = 0; i < V; ++i) acc[g][c][i] = fmaf(p, vf[c][i], acc[g][c][i] * corr);
    }
  }

  // ---- merge the warps of this CTA
#pragma unroll
  for (int g = 0; g < G; ++g) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int d = (c * 32 + lane) * V;
      if (d < hd)
#pragma unroll
        for (int i = 0; i < V; ++i) s_acc[((size_t)wid * G + g) * hd + d + i] = acc[g][c][i];
    }
    if (lane == 0) { s_m[wid * G + g] = m[g]; s_l[wid * G + g] = l[g]; }
  }
  __syncthreads();
  float* wsp = ws + (((size_t)b * nkv + kvh) * max_splits + split) * (size_t)(G * hd + 2 * G);
  for (int e = threadIdx.x; e < G * hd; e += DA_WARPS * 32) {
    const int g = e / hd;
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < DA_WARPS; ++w) M = fmaxf(M, s_m[w * G + g]);
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < DA_WARPS; ++w) {
      const float mw = s_m[w * G + g];
      if (mw != -INFINITY) a += __expf(mw - M) * s_acc[((size_t)w * G + g) * hd + (e % hd)];
    }
    wsp[e] = a;
  }
  if (threadIdx.x < G) {
    const int g = threadIdx.x;
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < DA_WARPS; ++w) M = fmaxf(M, s_m[w * G + g]);
    float L = 0.f;
#pragma unroll
    for (int w = 0; w < DA_WARPS; ++w) {
      const float mw = s_m[w * G + g];
      if (mw != -INFINITY) L += __expf(mw - M) * s_l[w * G + g];
    }
    wsp[G * hd + g] = M;
    wsp[G * hd + G + g] = L;
  }
  // ---- last CTA of this (b, kvh) merges the splits
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_ticket = atomicAdd(&counters[b * nkv + kvh], 1);
  __syncthreads();
  if (s_ticket != n_active - 1) return;
  __threadfence();
  const float* base = ws + ((size_t)b * nkv + kvh) * max_splits * (size_t)(G * hd + 2 * G);
  const size_t stride = (size_t)(G * hd + 2 * G);
  for (int e = threadIdx.x; e < G * hd; e += DA_WARPS * 32) {
    const int g = e / hd;
    float M = -INFINITY;
    for (int sidx = 0; sidx < n_active; ++sidx) M = fmaxf(M, __ldcg(base + sidx * stride + G * hd + g));
    float a = 0.f, L = 0.f;
    for (int sidx = 0; sidx < n_active; ++sidx) {
      const float w = __expf(__ldcg(base + sidx * stride + G * hd + g) - M);
      a += w * __ldcg(base + sidx * stride + e);
      L += w * __ldcg(base + sidx * stride + G * hd + G + g);
    }
    out[(size_t)b * nq * hd + (size_t)(kvh * G + g) * hd + (e % hd)] = from_f<T>(a / L);
  }
  if (threadIdx.x == 0) counters[b * nkv + kvh] = 0;  // ready for the next launch
}

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

long long pg_decode_attention_ws_floats(int B, int nq, int hd, int max_splits) {
  return (long long)B * max_splits * ((long long)nq * hd + 2LL * nq);
}

int pg_decode_attention(void* out, const void* q, const void* k_pool, const void* v_pool,
                        const int32_t* page_table, int pt_stride, int page_size, const int32_t* kv_len,
                        int kv_len_add, int B, int nq, int nkv, int hd, float scale_div, float* ws,
                        int* counters, int max_splits, int dtype, void* stream) {
  PG_REQUIRE(B > 0 && nq % nkv == 0 && max_splits >= 1, "decode_attention: bad shape");
  const int G = nq / nkv;
  dim3 grid(max_splits, B, nkv);
  const size_t smem = ((size_t)DA_WARPS * G * hd + 2 * DA_WARPS * G) * sizeof(float);
#define PG_DA_LAUNCH(GG, NCH)                                                                         \
  {                                                                                                   \
    auto kern = decode_attention_kernel<T, GG, NCH>;                                                  \
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    kern<<<grid, DA_WARPS * 32, smem, (cudaStream_t)stream>>>(                                        \
        (T*)out, (const T*)q, (const T*)k_pool, (const T*)v_pool, page_table, pt_stride, page_size,   \
        kv_len, kv_len_add, nq, nkv, hd, scale_div, ws, counters, max_splits);                        \
  }
  PG_DISPATCH_DTYPE(dtype, T, {
    constexpr int V = Vec<T>::N;
    PG_REQUIRE(hd % V == 0 && hd <= 64 * V, "decode_attention: unsupported head_dim %d", hd);
    const int nch = (hd + 32 * V - 1) / (32 * V);
    if (nch == 1) {
      switch (G) {
        case 1: PG_DA_LAUNCH(1, 1) break;
        case 2: PG_DA_LAUNCH(2, 1) break;
        case 4: PG_DA_LAUNCH(4, 1) break;
        case 8: PG_DA_LAUNCH(8, 1) break;
        default: set_error("decode_attention: unsupported group size %d", G); return PG_ERR_INVALID;
      }
    } else {
      switch (G) {
        case 1: PG_DA_LAUNCH(1, 2) break;
        case 2: PG_DA_LAUNCH(2, 2) break;
        case 4: PG_DA_LAUNCH(4, 2) break;
        case 8: PG_DA_LAUNCH(8, 2) break;
        default: set_error("decode_attention: unsupported group size %d", G); return PG_ERR_INVALID;
      }
    }
  });
#undef PG_DA_LAUNCH
  return check_launch("decode_attention");
}

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
