// Row / elementwise kernels of the PaliGemma path: embedding+merge, norms, im2col, RoPE+KV
// append (prefill), step bookkeeping, argmax, KV gather.  All HBM-bound; 128-bit accesses.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <atomic>

#include "common.cuh"
#include "tp_exchange.cuh"

namespace pg {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};
static thread_local Prefetch g_prefetch = {nullptr, 0};

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

Prefetch take_prefetch() {
  Prefetch p = g_prefetch;
  g_prefetch = {nullptr, 0};
  return p;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    cudaGetLastError();
    return PG_ERR_CUDA;
  }
  return PG_OK;
}

// ------------------------------------------------------------------ embed + merge
// One CTA per token.  The k-th image token (row-major over the flattened ids) takes image row k
// (masked_scatter order, modeling_gemma.py:498).
template <typename T>
__global__ void embed_merge_kernel(T* __restrict__ out, const int64_t* __restrict__ ids,
                                   const T* __restrict__ emb, const T* __restrict__ img,
                                   int n_tokens, int D, int64_t vocab, int64_t img_id, int64_t pad_id,
                                   int n_img_rows, float img_div, float normalizer, int* err_flag) {
  pdl_launch_dependents();  // a tensor-core kernel launched next may start its prologue now
  constexpr int V = Vec<T>::N;
  __shared__ float red[32];
  const int t = blockIdx.x;
  const int64_t id = ids[t];
  const T* src = nullptr;
  bool is_img = false;
  if (id == img_id) {
    float cnt = 0.f;
    for (int i = threadIdx.x; i < t; i += blockDim.x) cnt += (ids[i] == img_id) ? 1.f : 0.f;
    int k = (int)(block_sum(cnt, red) + 0.5f);
    if (img != nullptr && k < n_img_rows) { src = img + (size_t)k * D; is_img = true; }
    else if (err_flag && threadIdx.x == 0) *err_flag = 1;
  } else if (id != pad_id) {
    if (id >= 0 && id < vocab) src = emb + (size_t)id * D;
    else if (err_flag && threadIdx.x == 0) *err_flag = 1;
  }
  T* dst = out + (size_t)t * D;
  for (int c = threadIdx.x * V; c < D; c += blockDim.x * V) {
    float f[V];
    if (src) {
      unpack<T>(ldg_cached(src + c), f);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        float v = f[i];
        if (is_img) v = rnd<T>(v / img_div);
        f[i] = v * normalizer;
      }
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) f[i] = 0.f;
    }
    *reinterpret_cast<uint4*>(dst + c) = pack<T>(f);
  }
}

// ------------------------------------------------------------------ RMSNorm / LayerNorm
// One CTA per row; the row is kept in registers between the statistics pass and the write.
template <typename T, int MAXV>
__global__ void rmsnorm_kernel(T* __restrict__ out, const T* __restrict__ x, const T* __restrict__ w,
                               int D, float eps) {
  pdl_launch_dependents();  // a tensor-core kernel launched next may start its prologue now
  pdl_wait();               // (launched with the PDL attribute when the grid is small: inputs come from the predecessor)
  constexpr int V = Vec<T>::N;
  __shared__ float red[32];
  const T* xr = x + (size_t)blockIdx.x * D;
  T* orow = out + (size_t)blockIdx.x * D;
  float f[MAXV][V];
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < MAXV; ++j) {
    int c = (j * blockDim.x + threadIdx.x) * V;
    if (c < D) {
      unpack<T>(ldg_cached(xr + c), f[j]);
#pragma unroll
      for (int i = 0; i < V; ++i) ss += f[j][i] * f[j][i];
    }
  }
  ss = block_sum(ss, red);
  const float inv = rsqrtf(ss / (float)D + eps);
#pragma unroll
  for (int j = 0; j < MAXV; ++j) {
    int c = (j * blockDim.x + threadIdx.x) * V;
    if (c < D) {
      float g[V];
      unpack<T>(ldg_cached(w + c), g);
#pragma unroll
      for (int i = 0; i < V; ++i) f[j][i] = (f[j][i] * inv) * (1.0f + g[i]);
      *reinterpret_cast<uint4*>(orow + c) = pack<T>(f[j]);
    }
  }
}

template <typename T, int MAXV>
__global__ void layernorm_kernel(T* __restrict__ out, const T* __restrict__ x, const T* __restrict__ w,
                                 const T* __restrict__ b, int D, float eps) {
  pdl_launch_dependents();  // a tensor-core kernel launched next may start its prologue now
  pdl_wait();               // (launched with the PDL attribute when the grid is small: inputs come from the predecessor)
  constexpr int V = Vec<T>::N;
  __shared__ float red[32];
  const T* xr = x + (size_t)blockIdx.x * D;
  T* orow = out + (size_t)blockIdx.x * D;
  float f[MAXV][V];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < MAXV; ++j) {
    int c = (j * blockDim.x + threadIdx.x) * V;
    if (c < D) {
      unpack<T>(ldg_cached(xr + c), f[j]);
#pragma unroll
      for (int i = 0; i < V; ++i) s += f[j][i];
    }
  }
  const float mean = block_sum(s, red) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < MAXV; ++j) {
    int c = (j * blockDim.x + threadIdx.x) * V;
    if (c < D) {
#pragma unroll
      for (int i = 0; i < V; ++i) { float d = f[j][i] - mean; q += d * d; }
    }
  }
  const float rstd = rsqrtf(block_sum(q, red) / (float)D + eps);
#pragma unroll
  for (int j = 0; j < MAXV; ++j) {
    int c = (j * blockDim.x + threadIdx.x) * V;
    if (c < D) {
      float g[V], bb[V];
      unpack<T>(ldg_cached(w + c), g);
      unpack<T>(ldg_cached(b + c), bb);
#pragma unroll
      for (int i = 0; i < V; ++i) f[j][i] = (f[j][i] - mean) * rstd * g[i] + bb[i];
      *reinterpret_cast<uint4*>(orow + c) = pack<T>(f[j]);
    }
  }
}


// ------------------------------------------------------------------ warp-per-row norms (many rows)
// Prefill / vision: thousands of rows of 1152-2048 elements.  One warp per row (8 rows per CTA), the row
// stays in registers between the statistics and the write, fully coalesced 16-byte accesses.
template <typename T, int MAXV, bool LAYERNORM>
__global__ void __launch_bounds__(256)
norm_rows_warp_kernel(T* __restrict__ out, const T* __restrict__ x, const T* __restrict__ w, const T* __restrict__ b,
                      int rows, int D, float eps) {
  pdl_launch_dependents();  // a tensor-core kernel launched next may start its prologue now
  pdl_wait();               // (launched with the PDL attribute when the grid is small: inputs come from the predecessor)
  constexpr int V = Vec<T>::N;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const T* xr = x + (size_t)row * D;
  T* orow = out + (size_t)row * D;
  // the row is held as the raw 16-byte vectors (4 registers each, unpacked on the fly in every pass): ~45 registers,
  // so 5-6 CTAs (40+ rows) per SM keep enough bytes in flight for HBM
  uint4 raw[MAXV];
#pragma unroll
  for (int j = 0; j < MAXV; ++j) {
    const int c = (j * 32 + lane) * V;
    raw[j] = c < D ? ldg_cached(xr + c) : make_uint4(0, 0, 0, 0);
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < MAXV; ++j) {
    float f[V];
    unpack<T>(raw[j], f);   // out-of-row vectors are zeros: they add nothing to either sum
#pragma unroll
    for (int i = 0; i < V; ++i) s += LAYERNORM ? f[i] : f[i] * f[i];
  }
  s = warp_sum(s);
  float mean = 0.f, scale;
  if (LAYERNORM) {
    mean = s / (float)D;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
      const int c = (j * 32 + lane) * V;
      if (c < D) {
        float f[V];
        unpack<T>(raw[j], f);
#pragma unroll
        for (int i = 0; i < V; ++i) { const float d = f[i] - mean; q += d * d; }
      }
    }
    scale = rsqrtf(warp_sum(q) / (float)D + eps);
  } else {
    scale = rsqrtf(s / (float)D + eps);
  }
#pragma unroll
  for (int j = 0; j < MAXV; ++j) {
    const int c = (j * 32 + lane) * V;
    if (c < D) {
      float f[V], g[V], bb[V];
      unpack<T>(raw[j], f);
      unpack<T>(ldg_cached(w + c), g);
      if (LAYERNORM) unpack<T>(ldg_cached(b + c), bb);
#pragma unroll
      for (int i = 0; i < V; ++i)
        f[i] = LAYERNORM ? (f[i] - mean) * scale * g[i] + bb[i] : (f[i] * scale) * (1.0f + g[i]);
      *reinterpret_cast<uint4*>(orow + c) = pack<T>(f);
    }
  }
}

// ------------------------------------------------------------------ im2col (stride == kernel)
template <typename T>
__global__ void im2col_kernel(T* __restrict__ out, const T* __restrict__ px, int C, int H, int W, int p,
                              int ld_out, long long total) {
  pdl_launch_dependents();  // a tensor-core kernel launched next may start its prologue now
  const int G = W / p, P = (H / p) * G, Kc = C * p * p;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int col = (int)(i % ld_out);
    long long row = i / ld_out;
    T v = from_f<T>(0.f);
    if (col < Kc) {
      int b = (int)(row / P), pp = (int)(row % P);
      int py = pp / G, pxx = pp % G;
      int c = col / (p * p), r = col % (p * p);
      int ky = r / p, kx = r % p;
      v = px[(((size_t)b * C + c) * H + (py * p + ky)) * W + (pxx * p + kx)];
    }
    out[i] = v;
  }
}

// ------------------------------------------------------------------ RoPE + KV append (prefill)
// One CTA per token.  Thread j < hd/2 handles the rotation pair (j, j+hd/2) of every head.
template <typename T>
__global__ void rope_append_kernel(T* __restrict__ q_out, const T* __restrict__ qkv,
                                   const float* __restrict__ inv_freq, const int32_t* __restrict__ positions,
                                   T* __restrict__ k_pool, T* __restrict__ v_pool,
                                   const int32_t* __restrict__ page_table, int pt_stride, int page_size,
                                   const int32_t* __restrict__ slot_base, int q_len, int nq, int nkv, int hd,
                                   int max_pos) {
  pdl_launch_dependents();  // a tensor-core kernel launched next may start its prologue now
  pdl_wait();               // (launched with the PDL attribute when the grid is small: inputs come from the predecessor)
  const int t = blockIdx.x, b = t / q_len, i = t % q_len;
  const int half = hd / 2;
  int pos = positions[t];
  pos = min(max(pos, 0), max_pos - 1);
  const int slot = slot_base[b] + i;
  const int page = page_table[(size_t)b * pt_stride + slot / page_size];
  const size_t kv_row = ((size_t)page * page_size + (slot % page_size)) * (size_t)(nkv * hd);
  const T* row = qkv + (size_t)t * (nq + 2 * nkv) * hd;
  T* qo = q_out + (size_t)t * nq * hd;
  // a thread owns one rotation index j (cos / sin evaluated once, reused for every head) and every other head
  const int jl = threadIdx.x % (blockDim.x / 2), hg = threadIdx.x / (blockDim.x / 2);
  for (int j = jl; j < half; j += blockDim.x / 2) {
    const float ang = (float)pos * inv_freq[j];
    const float c = rnd<T>(cosf(ang)), s = rnd<T>(sinf(ang));
#pragma unroll 3
    for (int h = hg; h < nq + nkv; h += 2) {
      const float x1 = to_f<T>(row[h * hd + j]), x2 = to_f<T>(row[h * hd + j + half]);
      const float o1 = rnd<T>(rnd<T>(x1 * c) + rnd<T>(-x2 * s));
      const float o2 = rnd<T>(rnd<T>(x2 * c) + rnd<T>(x1 * s));
      if (h < nq) {
        qo[h * hd + j] = from_f<T>(o1);
        qo[h * hd + j + half] = from_f<T>(o2);
      } else {
        const int kh = h - nq;
        k_pool[kv_row + kh * hd + j] = from_f<T>(o1);
        k_pool[kv_row + kh * hd + j + half] = from_f<T>(o2);
      }
    }
  }
  const T* vrow = row + (size_t)(nq + nkv) * hd;
  for (int u = threadIdx.x; u < nkv * hd; u += blockDim.x) v_pool[kv_row + u] = vrow[u];
}

// ------------------------------------------------------------------ decode-step bookkeeping
__global__ void step_advance_kernel(int64_t* next_ids, int64_t* history, int hist_stride, int* step_counter,
                                    unsigned long long* keys, const int64_t* sampled, int32_t* kv_len,
                                    int32_t* positions, int B, TpEx ex) {
  const int b = threadIdx.x;
  const int step = step_counter ? *step_counter : 0;
  __syncthreads();
  if (b < B) {
    // tensor parallel + greedy: the winning (value, global index) key over the ranks' vocabulary shards
    int64_t tok = sampled ? sampled[b]
                          : (int64_t)argmax_key_index(ex.peers ? tp_wait_best_key(ex, b) : keys[b]);
    next_ids[b] = tok;
    // ring index: the step counter (also the sampler's RNG offset) keeps counting for the life of the engine
    if (history && hist_stride > 0) history[(size_t)b * hist_stride + (unsigned)step % (unsigned)hist_stride] = tok;
    if (kv_len) kv_len[b] += 1;
    if (positions) positions[b] += 1;
    keys[b] = 0ull;
  }
  if (b == 0 && step_counter) *step_counter = (int)((unsigned)step + 1u);   // wraps like the ring index above
}

// One launch at the head of an API-driven decode step: token ids and positions into the graph's static buffers,
// plus the reference's all-ones attention-mask check (modeling_gemma.py:559) reported through a flag instead of a
// host synchronisation.
template <typename M>
__device__ __forceinline__ bool mask_is_one(const void* mask, long long i) {
  return reinterpret_cast<const M*>(mask)[i] == (M)1;
}
__global__ void decode_inputs_kernel(int64_t* ids_dst, const int64_t* ids_src, int32_t* positions, int position,
                                     const void* mask, int mask_kind, long long mask_n, int* bad_flag, int B) {
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    ids_dst[b] = ids_src[b];
    positions[b] = position;
  }
  bool bad = false;
  for (long long i = threadIdx.x; i < mask_n; i += blockDim.x) {
    bool one;
    switch (mask_kind) {
      case 0: one = mask_is_one<int64_t>(mask, i); break;
      case 1: one = mask_is_one<float>(mask, i); break;
      case 2: one = mask_is_one<int32_t>(mask, i); break;
      case 3: one = reinterpret_cast<const uint16_t*>(mask)[i] == 0x3F80u; break;  // bf16 1.0
      case 4: one = reinterpret_cast<const uint16_t*>(mask)[i] == 0x3C00u; break;  // f16 1.0
      case 5: one = mask_is_one<uint8_t>(mask, i); break;
      default: one = mask_is_one<double>(mask, i); break;
    }
    bad |= !one;
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0 && bad_flag) {
    *bad_flag = 1;
    __threadfence_system();
  }
}

// ------------------------------------------------------------------ argmax over fp32 logits
__global__ void argmax_partial_kernel(unsigned long long* keys, const float* __restrict__ logits, long long V) {
  const int b = blockIdx.y;
  const float* row = logits + (size_t)b * V;
  unsigned long long best = 0ull;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < V;
       i += (long long)gridDim.x * blockDim.x) {
    unsigned long long k = argmax_key(row[i], (unsigned)i);
    best = k > best ? k : best;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
    best = other > best ? other : best;
  }
  if ((threadIdx.x & 31) == 0 && best) atomicMax(&keys[b], best);
}
__global__ void argmax_final_kernel(int64_t* out, unsigned long long* keys, int B) {
  int b = threadIdx.x;
  if (b < B) { out[b] = (int64_t)argmax_key_index(keys[b]); keys[b] = 0ull; }
}

// ------------------------------------------------------------------ KV gather
template <typename T>
__global__ void kv_gather_kernel(T* __restrict__ out, const T* __restrict__ pool,
                                 const int32_t* __restrict__ page_table, int pt_stride, int page_size,
                                 int T_len, int nkv, int hd) {
  // out: [B, nkv, T, hd]; grid (T, B)
  const int j = blockIdx.x, b = blockIdx.y;
  const int page = page_table[(size_t)b * pt_stride + j / page_size];
  const T* src = pool + ((size_t)page * page_size + (j % page_size)) * (size_t)(nkv * hd);
  for (int u = threadIdx.x; u < nkv * hd; u += blockDim.x) {
    int kh = u / hd, d = u % hd;
    out[(((size_t)b * nkv + kh) * T_len + j) * hd + d] = src[u];
  }
}

}  // namespace pg

using namespace pg;

extern "C" {

const char* pg_last_error(void) { return g_err; }
int pg_abi_version(void) { return 1; }
unsigned long long pg_launch_count(void) { return g_launches.load(); }

int pg_set_next_prefetch(const void* ptr, long long bytes) {
  PG_REQUIRE(bytes >= 0 && ((uintptr_t)ptr % 16) == 0, "set_next_prefetch: region must be 16-byte aligned");
  g_prefetch = {(const char*)ptr, (unsigned long long)(ptr ? bytes : 0)};
  return PG_OK;
}

int pg_embed_merge(void* out, const int64_t* ids, const void* emb, const void* img_feats, int n_tokens,
                   int D, int64_t vocab, int64_t image_token_id, int64_t pad_id, int n_img_rows,
                   float img_div, float normalizer, int* err_flag, int dtype, void* stream) {
  if (n_tokens <= 0) return PG_OK;
  PG_DISPATCH_DTYPE(dtype, T, {
    PG_REQUIRE(D % Vec<T>::N == 0, "embed_merge: D=%d not a multiple of %d", D, Vec<T>::N);
    embed_merge_kernel<T><<<n_tokens, 256, 0, (cudaStream_t)stream>>>(
        (T*)out, ids, (const T*)emb, (const T*)img_feats, n_tokens, D, vocab, image_token_id, pad_id,
        n_img_rows, img_div, normalizer, err_flag);
  });
  return check_launch("embed_merge");
}

int pg_rmsnorm(void* out, const void* x, const void* w, int rows, int D, float eps, int dtype, void* stream) {
  if (rows <= 0) return PG_OK;
  PG_DISPATCH_DTYPE(dtype, T, {
    constexpr int V = Vec<T>::N;
    PG_REQUIRE(D % V == 0 && D <= 256 * V * 4, "rmsnorm: unsupported D=%d", D);
    const T* none = nullptr;
    if (rows >= 64 && D <= 32 * V * 5)
      return launch_tc("rmsnorm", norm_rows_warp_kernel<T, 5, false>, dim3(cdiv(rows, 8)), dim3(256), 0, 1, rows <= 4096,
                       (cudaStream_t)stream, (T*)out, (const T*)x, (const T*)w, none, rows, D, eps);
    if (rows >= 64 && D <= 32 * V * 8)
      return launch_tc("rmsnorm", norm_rows_warp_kernel<T, 8, false>, dim3(cdiv(rows, 8)), dim3(256), 0, 1, rows <= 4096,
                       (cudaStream_t)stream, (T*)out, (const T*)x, (const T*)w, none, rows, D, eps);
    return launch_tc("rmsnorm", rmsnorm_kernel<T, 4>, dim3(rows), dim3(256), 0, 1, rows <= 1024, (cudaStream_t)stream, (T*)out,
                     (const T*)x, (const T*)w, D, eps);
  });
  return PG_OK;
}

int pg_layernorm(void* out, const void* x, const void* w, const void* b, int rows, int D, float eps,
                 int dtype, void* stream) {
  if (rows <= 0) return PG_OK;
  PG_DISPATCH_DTYPE(dtype, T, {
    constexpr int V = Vec<T>::N;
    PG_REQUIRE(D % V == 0 && D <= 256 * V * 4, "layernorm: unsupported D=%d", D);
    if (rows >= 64 && D <= 32 * V * 5)
      return launch_tc("layernorm", norm_rows_warp_kernel<T, 5, true>, dim3(cdiv(rows, 8)), dim3(256), 0, 1, rows <= 4096,
                       (cudaStream_t)stream, (T*)out, (const T*)x, (const T*)w, (const T*)b, rows, D, eps);
    if (rows >= 64 && D <= 32 * V * 8)
      return launch_tc("layernorm", norm_rows_warp_kernel<T, 8, true>, dim3(cdiv(rows, 8)), dim3(256), 0, 1, rows <= 4096,
                       (cudaStream_t)stream, (T*)out, (const T*)x, (const T*)w, (const T*)b, rows, D, eps);
    return launch_tc("layernorm", layernorm_kernel<T, 4>, dim3(rows), dim3(256), 0, 1, rows <= 1024, (cudaStream_t)stream,
                     (T*)out, (const T*)x, (const T*)w, (const T*)b, D, eps);
  });
  return PG_OK;
}

int pg_im2col(void* out, const void* pixels, int B, int C, int H, int W, int p, int ld_out, int dtype,
              void* stream) {
  PG_REQUIRE(H % p == 0 && W % p == 0 && ld_out >= C * p * p, "im2col: bad geometry");
  long long total = (long long)B * (H / p) * (W / p) * ld_out;
  if (total <= 0) return PG_OK;
  int grid = (int)((total + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  PG_DISPATCH_DTYPE(dtype, T, {
    im2col_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>((T*)out, (const T*)pixels, C, H, W, p, ld_out, total);
  });
  return check_launch("im2col");
}

int pg_rope_append(void* q_out, const void* qkv, const float* inv_freq, const int32_t* positions,
                   void* k_pool, void* v_pool, const int32_t* page_table, int pt_stride, int page_size,
                   const int32_t* slot_base, int B, int q_len, int nq, int nkv, int hd, int max_pos,
                   int dtype, void* stream) {
  if (B * q_len <= 0) return PG_OK;
  PG_REQUIRE(hd % 2 == 0, "rope_append: odd head_dim");
  PG_DISPATCH_DTYPE(dtype, T, {
    return launch_tc("rope_append", rope_append_kernel<T>, dim3(B * q_len), dim3(256), 0, 1, B * q_len <= 1024,
                     (cudaStream_t)stream, (T*)q_out, (const T*)qkv, inv_freq, positions, (T*)k_pool, (T*)v_pool, page_table,
                     pt_stride, page_size, slot_base, q_len, nq, nkv, hd, max_pos);
  });
  return PG_OK;
}

int pg_step_advance(int64_t* next_ids, int64_t* history, int hist_stride, int* step_counter,
                    unsigned long long* keys, const int64_t* sampled, int32_t* kv_len, int32_t* positions,
                    int B, const pg_tp_exchange* keys_ex, void* stream) {
  PG_REQUIRE(B > 0 && B <= 1024, "step_advance: B=%d", B);
  const TpEx tex = tp_ex_from(keys_ex);
  PG_REQUIRE(!keys_ex || (tex.tp >= 2 && tex.tp <= TP_MAX_RANKS && (long long)B * 16 <= tex.slot_bytes),
             "step_advance: bad tensor-parallel key exchange");
  step_advance_kernel<<<1, ((B + 31) / 32) * 32, 0, (cudaStream_t)stream>>>(
      next_ids, history, hist_stride, step_counter, keys, sampled, kv_len, positions, B, tex);
  return check_launch("step_advance");
}

int pg_decode_inputs(int64_t* ids_dst, const int64_t* ids_src, int32_t* positions, int position, const void* mask,
                     int mask_kind, long long mask_n, int* bad_flag, int B, void* stream) {
  PG_REQUIRE(B > 0 && ids_dst && ids_src && positions, "decode_inputs: bad arguments");
  PG_REQUIRE(!mask || (mask_kind >= 0 && mask_kind <= 6), "decode_inputs: unknown mask element kind %d", mask_kind);
  decode_inputs_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(ids_dst, ids_src, positions, position, mask, mask_kind,
                                                           mask ? mask_n : 0, bad_flag, B);
  return check_launch("decode_inputs");
}

int pg_argmax(int64_t* out, const float* logits, unsigned long long* keys, int B, int64_t V, void* stream) {
  PG_REQUIRE(B > 0 && B <= 1024, "argmax: B=%d", B);
  int gx = (int)((V + 256 * 8 - 1) / (256 * 8));
  if (gx > 148) gx = 148;
  argmax_partial_kernel<<<dim3(gx, B), 256, 0, (cudaStream_t)stream>>>(keys, logits, V);
  if (out) argmax_final_kernel<<<1, ((B + 31) / 32) * 32, 0, (cudaStream_t)stream>>>(out, keys, B);
  return check_launch("argmax");
}

int pg_kv_gather(void* out, const void* pool, const int32_t* page_table, int pt_stride, int page_size, int B,
                 int T_len, int nkv, int hd, int dtype, void* stream) {
  if (B * T_len <= 0) return PG_OK;
  PG_DISPATCH_DTYPE(dtype, T, {
    kv_gather_kernel<T><<<dim3(T_len, B), 128, 0, (cudaStream_t)stream>>>((T*)out, (const T*)pool, page_table,
                                                                          pt_stride, page_size, T_len, nkv, hd);
  });
  return check_launch("kv_gather");
}

}  // extern "C"
