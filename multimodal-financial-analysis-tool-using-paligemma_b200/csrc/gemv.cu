// Decode-path weight-streaming kernels (q_len == 1): every weight byte is read exactly once per
// step with 128-bit loads, fp32 accumulation, and the reference's elementwise neighbours fused
// in: RMSNorm prologue, RoPE + KV-append / GeGLU / residual / logits+argmax epilogues.
// HBM-bound: bytes per step = the weight bytes.
//
// Latency structure (what matters at batch 1, where a layer is five ~3-25 us kernels):
//  * a stage is U x R independent 128-bit loads per lane, 24-32 warps per SM keep > 100 KB in flight;
//  * programmatic dependent launch: a kernel's CTAs start while its predecessor drains, issue
//    their first weight loads and the L2 prefetch of the next kernel's weights, and only then
//    execute griddepcontrol.wait (weights never depend on activations);
#include <cstdlib>

#include "common.cuh"
#include "tp_exchange.cuh"

namespace pg {

#ifndef PG_GEMV_THREADS
#define PG_GEMV_THREADS 256
#endif
constexpr int GEMV_THREADS = PG_GEMV_THREADS;
constexpr int GEMV_WARPS = GEMV_THREADS / 32;

// One pipeline stage: U 128-bit vectors of each of R weight rows (this lane's share).
template <int R, int U>
struct WChunk {
  uint4 v[U][R];
};

template <typename T, int R, int U>
__device__ __forceinline__ void load_chunk(WChunk<R, U>& c, const T* const (&wrow)[R], int k0, int k_end) {
  constexpr int V = Vec<T>::N;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int k = k0 + u * 32 * V;
#pragma unroll
    for (int r = 0; r < R; ++r) c.v[u][r] = (k < k_end) ? ldg_stream(wrow[r] + k) : make_uint4(0, 0, 0, 0);
  }
}

// acc[r][b] += W[row r][k..] * x[b][k..] for one stage.  X_SMEM: x is fp32 in shared memory
// (normed prologue); else x is model-dtype global memory (L1-resident, a few KB).
template <typename T, int NB, int R, int U, bool X_SMEM>
__device__ __forceinline__ void consume_chunk(const WChunk<R, U>& c, const float* __restrict__ xs,
                                              const T* __restrict__ xg, int K, int k0, int k_end,
                                              float (&acc)[R][NB]) {
  constexpr int V = Vec<T>::N;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int k = k0 + u * 32 * V;
    if (k < k_end) {
      float wf[R][V];
#pragma unroll
      for (int r = 0; r < R; ++r) unpack<T>(c.v[u][r], wf[r]);
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        float xf[V];
        if (X_SMEM) {
#pragma unroll
          for (int i = 0; i < V; i += 4) {
            float4 t = *reinterpret_cast<const float4*>(xs + (size_t)b * K + k + i);
            xf[i] = t.x; xf[i + 1] = t.y; xf[i + 2] = t.z; xf[i + 3] = t.w;
          }
        } else {
          unpack<T>(ldg_cached(xg + (size_t)b * K + k), xf);
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int i = 0; i < V; ++i) acc[r][b] = fmaf(wf[r][i], xf[i], acc[r][b]);
      }
    }
  }
}

// Streams the rows of a sequence of "units" through one warp.  rows(u, wrow) fills the R row
// pointers of unit u; done(u, acc) receives the warp-reduced sums.  prime() (first stage of the
// warp's first unit) is called before the kernel's dependency wait, run() after it.  A stage is
// U x R independent 128-bit loads per lane; memory-level parallelism beyond that comes from
// occupancy (3-4 CTAs x 8 warps per SM), which measured faster than register double-buffering.
template <typename T, int NB, int R, int U, bool X_SMEM>
struct RowStreamer {
  static constexpr int V = Vec<T>::N;
  static constexpr int STEP = 32 * V * U;
  WChunk<R, U> cur;
  const T* wrow[R];
  long long u;  // current unit

  template <typename RowsFn>
  __device__ __forceinline__ void prime(long long first, long long n_units, int k_begin, int k_end, RowsFn rows) {
    u = first;
    if (u < n_units) {
      rows(u, wrow);
      load_chunk<T, R, U>(cur, wrow, k_begin + (int)(threadIdx.x & 31) * V, k_end);
    }
  }

  template <typename RowsFn, typename DoneFn>
  __device__ __forceinline__ void run(long long n_units, long long stride, const float* xs, const T* xg, int K,
                                      int k_begin, int k_end, RowsFn rows, DoneFn done) {
    run(n_units, stride, xs, xg, K, k_begin, k_end, rows, done, [](long long) {});
  }

  // pre(u) runs before unit u's rows are consumed: epilogue inputs that do not depend on the dot products (addresses,
  // angles) are fetched / computed while the weight loads are in flight instead of after the reduction
  template <typename RowsFn, typename DoneFn, typename PreFn>
  __device__ __forceinline__ void run(long long n_units, long long stride, const float* xs, const T* xg, int K,
                                      int k_begin, int k_end, RowsFn rows, DoneFn done, PreFn pre) {
    const int lane_off = (int)(threadIdx.x & 31) * V;
    bool primed = true;  // the first stage of the first unit was loaded by prime()
    while (u < n_units) {
      pre(u);
      float acc[R][NB];
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int b = 0; b < NB; ++b) acc[r][b] = 0.f;
      // base is warp-uniform, so every lane runs the same number of stages
      for (int base = k_begin; base < k_end; base += STEP) {
        if (!primed) load_chunk<T, R, U>(cur, wrow, base + lane_off, k_end);
        primed = false;
        consume_chunk<T, NB, R, U, X_SMEM>(cur, xs, xg, K, base + lane_off, k_end, acc);
      }
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int b = 0; b < NB; ++b) acc[r][b] = warp_sum(acc[r][b]);
      done(u, acc);
      u += stride;
      if (u < n_units) rows(u, wrow);
    }
  }
};

// RMSNorm of NB rows into shared memory as fp32 values already rounded to the model dtype
// (the reference materialises the normed tensor: modeling_gemma.py:120).
template <typename T, int NB>
__device__ __forceinline__ void norm_rows_to_smem(float* xs, const T* __restrict__ x, const T* __restrict__ w,
                                                  int D, float eps) {
  constexpr int V = Vec<T>::N;
  __shared__ float red[NB][GEMV_WARPS];
  float ss[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) ss[b] = 0.f;
  for (int c = threadIdx.x * V; c < D; c += GEMV_THREADS * V) {
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      float f[V];
      unpack<T>(ldg_cached(x + (size_t)b * D + c), f);
#pragma unroll
      for (int i = 0; i < V; ++i) ss[b] += f[i] * f[i];
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    float v = warp_sum(ss[b]);
    if (lane == 0) red[b][wid] = v;
  }
  __syncthreads();
  float inv[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < GEMV_WARPS; ++i) t += red[b][i];
    inv[b] = rsqrtf(t / (float)D + eps);
  }
  for (int c = threadIdx.x * V; c < D; c += GEMV_THREADS * V) {
    float g[V];
    unpack<T>(ldg_cached(w + c), g);
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      float f[V];
      unpack<T>(ldg_cached(x + (size_t)b * D + c), f);
#pragma unroll
      for (int i = 0; i < V; ++i) xs[(size_t)b * D + c + i] = rnd<T>((f[i] * inv[b]) * (1.0f + g[i]));
    }
  }
  __syncthreads();
}

// Tensor-parallel variant of the prologue (tp_exchange.cuh): the input row is not in memory yet -- its last term
// is still spread over the ranks as fp32 partials of the previous o_proj / down_proj.  Every CTA waits for the
// partials in its local exchange buffer, sums them in rank order, rounds once (the reference's projection output),
// adds the residual stream x_in (rounded again: `residual + hidden_states`), keeps the new stream in shared memory
// and normalises it there.  CTA 0 also writes the new stream to x_out for the next residual add.
template <typename T, int NB>
__device__ __forceinline__ void reduce_norm_rows_to_smem(float* xs, const T* __restrict__ x_in, T* __restrict__ x_out,
                                                         const T* __restrict__ w, int D, float eps, const TpEx& ex) {
  __shared__ float red_tp[NB][GEMV_WARPS];
  const uint32_t seq = tp_seq(ex);
  bool dead = tp_failed(ex);
  float ss[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) ss[b] = 0.f;
  const int half = D / 2;
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    for (int c2 = threadIdx.x; c2 < half; c2 += GEMV_THREADS) {
      const size_t i = (size_t)b * D + 2 * c2;
      const float2 p = tp_reduce_pair(ex, seq, (long long)(i >> 1), dead);
      const float v0 = rnd<T>(to_f<T>(x_in[i]) + rnd<T>(p.x));
      const float v1 = rnd<T>(to_f<T>(x_in[i + 1]) + rnd<T>(p.y));
      xs[i] = v0;
      xs[i + 1] = v1;
      ss[b] += v0 * v0 + v1 * v1;
      if (blockIdx.x == 0 && x_out) {
        x_out[i] = from_f<T>(v0);
        x_out[i + 1] = from_f<T>(v1);
      }
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    float v = warp_sum(ss[b]);
    if (lane == 0) red_tp[b][wid] = v;
  }
  __syncthreads();
  float inv[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < GEMV_WARPS; ++i) t += red_tp[b][i];
    inv[b] = rsqrtf(t / (float)D + eps);
  }
  for (int c = threadIdx.x; c < D; c += GEMV_THREADS) {
    const float g = 1.0f + to_f<T>(w[c]);
#pragma unroll
    for (int b = 0; b < NB; ++b) xs[(size_t)b * D + c] = rnd<T>((xs[(size_t)b * D + c] * inv[b]) * g);
  }
  __syncthreads();
}

// TPX is a template parameter so the single-GPU instantiations keep their register budget (72 registers for the
// gate/up kernel = 3 CTAs per SM; the exchange code would push it past 100)
template <typename T, int NB, bool TPX>
__device__ __forceinline__ void prologue_rows(float* xs, const T* __restrict__ x, T* __restrict__ x_out,
                                              const T* __restrict__ w, int D, float eps, const TpEx& ex) {
  if constexpr (TPX) reduce_norm_rows_to_smem<T, NB>(xs, x, x_out, w, D, eps, ex);
  else norm_rows_to_smem<T, NB>(xs, x, w, D, eps);
}

// ------------------------------------------------------------------ RMSNorm + QKV + RoPE + KV append
template <typename T, int NB, int QKV_U, bool TPX>
__global__ void __launch_bounds__(GEMV_THREADS)
decode_qkv_kernel(T* __restrict__ q_out, const T* __restrict__ x, const T* __restrict__ norm_w,
                  const T* __restrict__ W, const float* __restrict__ inv_freq,
                  const int32_t* __restrict__ positions, T* __restrict__ k_pool, T* __restrict__ v_pool,
                  const int32_t* __restrict__ page_table, int pt_stride, int page_size,
                  const int32_t* __restrict__ kv_len, int D, int nq, int nkv, int hd, float eps, int max_pos,
                  Prefetch pf, TpEx ex, T* __restrict__ x_out) {
  extern __shared__ __align__(16) float xs[];
  pdl_launch_dependents();
  l2_prefetch_slice(pf);
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * GEMV_WARPS + (threadIdx.x >> 5), nwarps = gridDim.x * GEMV_WARPS;
  const int half = hd / 2, units = (nq + 2 * nkv) * half;
  auto rows = [&](long long u, const T* (&wr)[2]) {
    const int h = (int)u / half, j = (int)u % half;
    wr[0] = W + (size_t)(h * hd + j) * D;
    wr[1] = W + (size_t)(h * hd + j + half) * D;
  };
  RowStreamer<T, NB, 2, QKV_U, true> rs;
  rs.prime(warp, units, 0, D, rows);
  pdl_wait();
  // lane b < NB owns batch row b in the epilogue: its cache slot and position do not depend on the projections, so
  // they are fetched here (the loads overlap the norm prologue) and the RoPE angle of a unit is computed in pre()
  size_t kv_row = 0;
  int pos = 0;
  if (lane < NB) {
    const int slot = kv_len[lane];
    const int page = page_table[(size_t)lane * pt_stride + slot / page_size];
    kv_row = ((size_t)page * page_size + (slot % page_size)) * (size_t)(nkv * hd);
    pos = min(max(positions[lane], 0), max_pos - 1);
  }
  prologue_rows<T, NB, TPX>(xs, x, x_out, norm_w, D, eps, ex);
  float rc = 1.f, rsn = 0.f;
  rs.run(units, nwarps, xs, nullptr, D, 0, D, rows, [&](long long u, float (&acc)[2][NB]) {
    if (lane < NB) {
      const int h = (int)u / half, j = (int)u % half;
      float a1 = 0.f, a2 = 0.f;
#pragma unroll
      for (int b = 0; b < NB; ++b) if (b == lane) { a1 = acc[0][b]; a2 = acc[1][b]; }
      const int b = lane;
      const float x1 = rnd<T>(a1), x2 = rnd<T>(a2);
      if (h < nq + nkv) {
        const float c = rc, s = rsn;
        const float o1 = rnd<T>(rnd<T>(x1 * c) + rnd<T>(-x2 * s));
        const float o2 = rnd<T>(rnd<T>(x2 * c) + rnd<T>(x1 * s));
        if (h < nq) {
          T* qo = q_out + (size_t)b * nq * hd + h * hd;
          qo[j] = from_f<T>(o1);
          qo[j + half] = from_f<T>(o2);
        } else {
          T* ko = k_pool + kv_row + (h - nq) * hd;
          ko[j] = from_f<T>(o1);
          ko[j + half] = from_f<T>(o2);
        }
      } else {
        T* vo = v_pool + kv_row + (h - nq - nkv) * hd;
        vo[j] = from_f<T>(x1);
        vo[j + half] = from_f<T>(x2);
      }
    }
  }, [&](long long u) {
    if (lane < NB && (int)u / half < nq + nkv) {
      const float ang = (float)pos * inv_freq[(int)u % half];
      rc = rnd<T>(cosf(ang));
      rsn = rnd<T>(sinf(ang));
    }
  });
}

// ------------------------------------------------------------------ GEMV + residual (o_proj / down_proj)
// KS warps share one output row (split along K, combined through shared memory).
template <typename T, int NB, int KS, bool TPX>
__global__ void __launch_bounds__(GEMV_THREADS)
gemv_res_kernel(T* __restrict__ out, const T* __restrict__ x, const T* __restrict__ W, const T* __restrict__ R,
                int N, int K, Prefetch pf, TpEx ex) {
  constexpr int V = Vec<T>::N;
  constexpr int ROWS_PER_CTA = GEMV_WARPS / KS;
  __shared__ float part[2][GEMV_WARPS][NB];
  pdl_launch_dependents();
  l2_prefetch_slice(pf);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int rsub = wid / KS, ks = wid % KS;
  const int seg = ((K / V + KS - 1) / KS) * V;  // K-range of this warp, vector aligned
  const int k_begin = ks * seg, k_end = min(K, k_begin + seg);
  const long long n_iter = (N + ROWS_PER_CTA - 1) / ROWS_PER_CTA;
  // a row index past N is clamped for the loads and masked at the store
  auto rows = [&](long long it, const T* (&wr)[1]) {
    long long n = it * ROWS_PER_CTA + rsub;
    wr[0] = W + (size_t)(n < N ? n : N - 1) * K;
  };
  RowStreamer<T, NB, 1, 8, false> rs;
  rs.prime(blockIdx.x, n_iter, k_begin, k_end, rows);
  pdl_wait();
  // tensor parallel: this rank's fp32 partial goes to every rank's exchange buffer (this rank's own included)
  // instead of `out`; the next kernel's prologue sums the partials and adds the residual
  uint32_t seq = 0u;
  if constexpr (TPX) seq = tp_seq(ex);
  int parity = 0;
  rs.run(n_iter, gridDim.x, nullptr, x, K, k_begin, k_end, rows, [&](long long it, float (&acc)[1][NB]) {
    const long long n = it * ROWS_PER_CTA + rsub;
    if (KS > 1) {
      if (lane == 0) {
#pragma unroll
        for (int b = 0; b < NB; ++b) part[parity][wid][b] = acc[0][b];
      }
      __syncthreads();  // all warps of the CTA run the same iterations
      if (ks == 0) {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          float t = 0.f;
#pragma unroll
          for (int i = 0; i < KS; ++i) t += part[parity][rsub * KS + i][b];
          acc[0][b] = t;
        }
      }
      parity ^= 1;  // double-buffered: the next iteration's writes cannot race these reads
    }
    if constexpr (TPX) {
      // every lane of the ks == 0 warp holds the row's sums: lane (b, p) stores batch row b into rank p's buffer, so
      // the tp peer stores of a row are ONE instruction (lane 0 issuing them one after the other cost ~1.7 us at tp 8)
      if (ks == 0 && n < N) {
        const int tpn = ex.tp;
        if (NB * tpn <= 32) {
          if (lane < NB * tpn) {
            const int bb = lane / tpn, p = lane % tpn;
            float a = 0.f;
#pragma unroll
            for (int b = 0; b < NB; ++b) if (b == bb) a = acc[0][b];
            tp_store_word(tp_slot(ex, p, seq, ex.rank), (long long)bb * N + n, a, seq);
          }
        } else if (lane < NB) {
          float a = 0.f;
#pragma unroll
          for (int b = 0; b < NB; ++b) if (b == lane) a = acc[0][b];
          for (int p = 0; p < tpn; ++p) tp_store_word(tp_slot(ex, p, seq, ex.rank), (long long)lane * N + n, a, seq);
        }
      }
    } else if (ks == 0 && n < N && lane < NB) {
      float a = 0.f;
#pragma unroll
      for (int b = 0; b < NB; ++b) if (b == lane) a = acc[0][b];
      float v = rnd<T>(a);
      if (R) v = rnd<T>(to_f<T>(R[(size_t)lane * N + n]) + v);
      out[(size_t)lane * N + n] = from_f<T>(v);
    }
  });
}

// ------------------------------------------------------------------ RMSNorm + gate/up + GeGLU
template <typename T, int NB, bool TPX>
__global__ void __launch_bounds__(GEMV_THREADS)
decode_gateup_kernel(T* __restrict__ out, const T* __restrict__ x, const T* __restrict__ norm_w,
                     const T* __restrict__ W, int D, int F, float eps, Prefetch pf, TpEx ex, T* __restrict__ x_out) {
  extern __shared__ __align__(16) float xs[];
  pdl_launch_dependents();
  l2_prefetch_slice(pf);
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * GEMV_WARPS + (threadIdx.x >> 5), nwarps = gridDim.x * GEMV_WARPS;
  auto rows = [&](long long f, const T* (&wr)[2]) {
    wr[0] = W + (size_t)f * D;
    wr[1] = W + (size_t)(F + f) * D;
  };
  RowStreamer<T, NB, 2, 4, true> rs;
  rs.prime(warp, F, 0, D, rows);
  pdl_wait();
  prologue_rows<T, NB, TPX>(xs, x, x_out, norm_w, D, eps, ex);
  rs.run(F, nwarps, xs, nullptr, D, 0, D, rows, [&](long long f, float (&acc)[2][NB]) {
    if (lane < NB) {
      float g = 0.f, u = 0.f;
#pragma unroll
      for (int b = 0; b < NB; ++b) if (b == lane) { g = acc[0][b]; u = acc[1][b]; }
      const float act = rnd<T>(gelu_tanh(rnd<T>(g)));
      out[(size_t)lane * F + f] = from_f<T>(act * rnd<T>(u));
    }
  });
}

// ------------------------------------------------------------------ final RMSNorm + lm_head + argmax
template <typename T, int NB, bool TPX>
__global__ void __launch_bounds__(GEMV_THREADS)
decode_lmhead_kernel(float* __restrict__ logits, const T* __restrict__ x, const T* __restrict__ norm_w,
                     const T* __restrict__ W, int D, long long V, float eps, unsigned long long* keys,
                     Prefetch pf, TpEx ex, T* __restrict__ x_out) {
  extern __shared__ __align__(16) float xs[];
  __shared__ unsigned long long best_s[GEMV_WARPS][NB];
  pdl_launch_dependents();
  l2_prefetch_slice(pf);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long warp = (long long)blockIdx.x * GEMV_WARPS + wid, nwarps = (long long)gridDim.x * GEMV_WARPS;
  const long long units = (V + 1) / 2;
  auto rows = [&](long long u, const T* (&wr)[2]) {
    wr[0] = W + (size_t)(2 * u) * D;
    wr[1] = W + (size_t)((2 * u + 1 < V) ? 2 * u + 1 : 2 * u) * D;
  };
  RowStreamer<T, NB, 2, 4, true> rs;
  rs.prime(warp, units, 0, D, rows);
  pdl_wait();
  prologue_rows<T, NB, TPX>(xs, x, x_out, norm_w, D, eps, ex);
  unsigned long long best = 0ull;  // lane b < NB tracks batch row b
  rs.run(units, nwarps, xs, nullptr, D, 0, D, rows, [&](long long u, float (&acc)[2][NB]) {
    if (lane < NB) {
      const long long r0 = 2 * u, r1 = 2 * u + 1;
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int b = 0; b < NB; ++b) if (b == lane) { a0 = acc[0][b]; a1 = acc[1][b]; }
      a0 = rnd<T>(a0); a1 = rnd<T>(a1);
      float* lrow = logits + (size_t)lane * V;
      lrow[r0] = a0;
      unsigned long long k0 = argmax_key(a0, (unsigned)r0);
      best = k0 > best ? k0 : best;
      if (r1 < V) {
        lrow[r1] = a1;
        unsigned long long k1 = argmax_key(a1, (unsigned)r1);
        best = k1 > best ? k1 : best;
      }
    }
  });
  if (keys) {
    if (lane < NB) best_s[wid][lane] = best;
    __syncthreads();
    if (wid == 0 && lane < NB) {
      unsigned long long m = 0ull;
#pragma unroll
      for (int i = 0; i < GEMV_WARPS; ++i) m = best_s[i][lane] > m ? best_s[i][lane] : m;
      if (m) atomicMax(&keys[lane], m);
    }
  }
}

template <typename F>
static int dispatch_nb(int B, F&& f) {
  switch (B) {
    case 1: return f(std::integral_constant<int, 1>());
    case 2: return f(std::integral_constant<int, 2>());
    case 3: return f(std::integral_constant<int, 3>());
    case 4: return f(std::integral_constant<int, 4>());
    case 5: return f(std::integral_constant<int, 5>());
    case 6: return f(std::integral_constant<int, 6>());
    case 7: return f(std::integral_constant<int, 7>());
    case 8: return f(std::integral_constant<int, 8>());
  }
  set_error("decode batch %d outside [1,%d]", B, PG_MAX_DECODE_BATCH);
  return PG_ERR_INVALID;
}

// Persistent-style grids: a multiple of the 148 SMs.  Defaults measured on B200 (tools/kernel_sweep.py):
// 3 CTAs/SM for the RMSNorm-prologue kernels (72-register, 8 KB smem), 4 for the plain GEMV.
static int grid_for_units(long long units, int per_cta, bool norm_kernel, bool tp = false) {
  long long g = (units + per_cta - 1) / per_cta;
  static const int cps_norm = env_int("PG_CPS_NORM", 3), cps_res = env_int("PG_CPS_RES", 6);
  static const int cps_tp = env_int("PG_CPS_TP", 2);   // the exchange prologue costs ~110 registers: 2 CTAs per SM
  const long long cap = 148LL * (norm_kernel ? (tp ? cps_tp : cps_norm) : cps_res);
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace pg

using namespace pg;

#define PG_GR(KS_)                                                                                        \
  if (ex)                                                                                                 \
    return launch_pdl("gemv_res", gemv_res_kernel<T, NB, KS_, true>,                                      \
                      dim3(grid_for_units(cdiv(N, GEMV_WARPS / KS_), 1, false)), dim3(GEMV_THREADS), 0, st,   \
                      (T*)out, (const T*)x, (const T*)W, (const T*)R, N, K, pf, tex);                     \
  return launch_pdl("gemv_res", gemv_res_kernel<T, NB, KS_, false>,                                       \
                    dim3(grid_for_units(cdiv(N, GEMV_WARPS / KS_), 1, false)), dim3(GEMV_THREADS), 0, st,     \
                    (T*)out, (const T*)x, (const T*)W, (const T*)R, N, K, pf, tex)

#define PG_QKV(U_, TPX_)                                                                                          \
  return launch_pdl("decode_qkv", decode_qkv_kernel<T, NB, U_, TPX_>, dim3(grid_for_units(units, GEMV_WARPS, true, TPX_)), \
                    dim3(GEMV_THREADS), smem, (cudaStream_t)stream, (T*)q_out, (const T*)x, (const T*)norm_w,         \
                    (const T*)w_qkv, inv_freq, positions, (T*)k_pool, (T*)v_pool, page_table, pt_stride, page_size,  \
                    kv_len, D, nq, nkv, hd, eps, max_pos, pf, tex, (T*)x_out)

extern "C" {

int pg_decode_qkv(void* q_out, const void* x, const void* norm_w, const void* w_qkv, const float* inv_freq,
                  const int32_t* positions, void* k_pool, void* v_pool, const int32_t* page_table,
                  int pt_stride, int page_size, const int32_t* kv_len, int B, int D, int nq, int nkv, int hd,
                  float eps, int max_pos, const pg_tp_exchange* ex, void* x_out, int dtype, void* stream) {
  const Prefetch pf = take_prefetch();
  const TpEx tex = tp_ex_from(ex);
  PG_REQUIRE(!ex || (tex.tp >= 2 && tex.tp <= TP_MAX_RANKS && D % 2 == 0 && (long long)B * D * 8 <= tex.slot_bytes),
             "decode_qkv: bad tensor-parallel exchange");
  PG_REQUIRE(hd % 2 == 0, "decode_qkv: odd head_dim");
  const int units = (nq + 2 * nkv) * (hd / 2);
  PG_DISPATCH_DTYPE(dtype, T, {
    PG_REQUIRE(D % Vec<T>::N == 0, "decode_qkv: D=%d not vector aligned", D);
    return dispatch_nb(B, [&](auto nb) {
      constexpr int NB = decltype(nb)::value;
      size_t smem = (size_t)NB * D * sizeof(float);
      PG_REQUIRE(smem <= 200 * 1024, "decode_qkv: B*D too large for shared memory");
      static const int u8 = env_int("PG_QKV_U8", 1);
      if (ex) { PG_QKV(4, true); }
      if (NB <= 2 && u8) { PG_QKV(8, false); }
      PG_QKV(4, false);
    });
  });
  return PG_OK;
}

int pg_gemv_res(void* out, const void* x, const void* W, const void* R, int B, int N, int K,
                const pg_tp_exchange* ex, int dtype, void* stream) {
  const Prefetch pf = take_prefetch();
  const TpEx tex = tp_ex_from(ex);
  PG_REQUIRE(!ex || (tex.tp >= 2 && tex.tp <= TP_MAX_RANKS && (long long)B * N * 8 <= tex.slot_bytes),
             "gemv_res: bad tensor-parallel exchange");
  PG_DISPATCH_DTYPE(dtype, T, {
    PG_REQUIRE(K % Vec<T>::N == 0, "gemv_res: K=%d not vector aligned", K);
    return dispatch_nb(B, [&](auto nb) {
      constexpr int NB = decltype(nb)::value;
      static const int ks_big = env_int("PG_DOWN_KS", 4);
      cudaStream_t st = (cudaStream_t)stream;
      if (K >= 8192 && ks_big == 8) { PG_GR(8); }
      if (K >= 8192 && ks_big == 4) { PG_GR(4); }
      if (K >= 8192 && ks_big == 2) { PG_GR(2); }
      PG_GR(1);
    });
  });
  return PG_OK;
}

int pg_decode_gateup(void* out, const void* x, const void* norm_w, const void* w_gu, int B, int D, int F,
                     float eps, const pg_tp_exchange* ex, void* x_out, int dtype, void* stream) {
  const Prefetch pf = take_prefetch();
  const TpEx tex = tp_ex_from(ex);
  PG_REQUIRE(!ex || (tex.tp >= 2 && tex.tp <= TP_MAX_RANKS && D % 2 == 0 && (long long)B * D * 8 <= tex.slot_bytes),
             "decode_gateup: bad tensor-parallel exchange");
  PG_DISPATCH_DTYPE(dtype, T, {
    PG_REQUIRE(D % Vec<T>::N == 0, "decode_gateup: D=%d not vector aligned", D);
    return dispatch_nb(B, [&](auto nb) {
      constexpr int NB = decltype(nb)::value;
      size_t smem = (size_t)NB * D * sizeof(float);
      PG_REQUIRE(smem <= 200 * 1024, "decode_gateup: B*D too large for shared memory");
      if (ex)
        return launch_pdl("decode_gateup", decode_gateup_kernel<T, NB, true>, dim3(grid_for_units(F, GEMV_WARPS, true, true)),
                          dim3(GEMV_THREADS), smem, (cudaStream_t)stream, (T*)out, (const T*)x, (const T*)norm_w,
                          (const T*)w_gu, D, F, eps, pf, tex, (T*)x_out);
      return launch_pdl("decode_gateup", decode_gateup_kernel<T, NB, false>, dim3(grid_for_units(F, GEMV_WARPS, true)),
                        dim3(GEMV_THREADS), smem, (cudaStream_t)stream, (T*)out, (const T*)x, (const T*)norm_w,
                        (const T*)w_gu, D, F, eps, pf, tex, (T*)x_out);
    });
  });
  return PG_OK;
}

int pg_decode_lmhead(float* logits, const void* x, const void* norm_w, const void* w_emb, int B, int D,
                     int64_t V, float eps, unsigned long long* argmax_keys, const pg_tp_exchange* ex, void* x_out,
                     int dtype, void* stream) {
  const Prefetch pf = take_prefetch();
  const TpEx tex = tp_ex_from(ex);
  PG_REQUIRE(!ex || (tex.tp >= 2 && tex.tp <= TP_MAX_RANKS && D % 2 == 0 && (long long)B * D * 8 <= tex.slot_bytes),
             "decode_lmhead: bad tensor-parallel exchange");
  PG_DISPATCH_DTYPE(dtype, T, {
    PG_REQUIRE(D % Vec<T>::N == 0, "decode_lmhead: D=%d not vector aligned", D);
    return dispatch_nb(B, [&](auto nb) {
      constexpr int NB = decltype(nb)::value;
      size_t smem = (size_t)NB * D * sizeof(float);
      PG_REQUIRE(smem <= 200 * 1024, "decode_lmhead: B*D too large for shared memory");
      if (ex)
        return launch_pdl("decode_lmhead", decode_lmhead_kernel<T, NB, true>,
                          dim3(grid_for_units((V + 1) / 2, GEMV_WARPS, true, true)), dim3(GEMV_THREADS), smem,
                          (cudaStream_t)stream, logits, (const T*)x, (const T*)norm_w, (const T*)w_emb, D,
                          (long long)V, eps, argmax_keys, pf, tex, (T*)x_out);
      return launch_pdl("decode_lmhead", decode_lmhead_kernel<T, NB, false>,
                        dim3(grid_for_units((V + 1) / 2, GEMV_WARPS, true)), dim3(GEMV_THREADS), smem,
                        (cudaStream_t)stream, logits, (const T*)x, (const T*)norm_w, (const T*)w_emb, D,
                        (long long)V, eps, argmax_keys, pf, tex, (T*)x_out);
    });
  });
  return PG_OK;
}

}  // extern "C"
