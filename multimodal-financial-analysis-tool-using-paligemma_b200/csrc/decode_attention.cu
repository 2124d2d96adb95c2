// Decode attention: one new token per sequence against the paged KV cache
// (reference GemmaAttention.forward q_len==1, modeling_gemma.py:262-288; mask all zeros, fp32
// softmax).  The G query heads of a KV head share every K/V load (repeat_kv, :136-141, is never
// materialised).
//
// One thread-block cluster of DC_NS CTAs per (sequence, kv head): the cluster's DC_NS*8 warps
// take the cached tokens round-robin, each warp loads K and V rows for a batch of tokens at once
// (all loads in flight before the first use), keeps a flash-style (m, l, acc) per head in
// registers, the warps of a CTA merge through shared memory and the CTAs of the cluster merge
// through distributed shared memory — no global workspace, no atomics, one launch.
// Work is tiny (T*1 KB per layer); the design goal is latency, not bandwidth.
#include <cmath>
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace pg {

constexpr int DC_WARPS = 8;     // warps per CTA
constexpr int DC_MAX_PT = 1024; // page-table entries staged in shared memory

template <typename T, int G, int NCH, int TB, int DC_NS>
__global__ void __launch_bounds__(DC_WARPS * 32, 1)
decode_attention_cluster_kernel(T* __restrict__ out, const T* __restrict__ q, const T* __restrict__ k_pool,
                                const T* __restrict__ v_pool, const int32_t* __restrict__ page_table,
                                int pt_stride, int page_size, const int32_t* __restrict__ kv_len,
                                int kv_len_add, int nq, int nkv, int hd, float scale_div, float scale_mul,
                                Prefetch pf) {
  constexpr int V = Vec<T>::N;
  cg::cluster_group cluster = cg::this_cluster();
  pdl_launch_dependents();
  l2_prefetch_slice(pf);
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.y, kvh = blockIdx.z;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  pdl_wait();  // q and this step's K/V row come from the preceding qkv kernel
  const int T_len = kv_len[b] + kv_len_add;
  const int row_elems = nkv * hd;

  extern __shared__ __align__(16) float sm[];
  float* s_acc = sm;                                     // [DC_WARPS][G][hd]
  float* s_m = s_acc + (size_t)DC_WARPS * G * hd;        // [DC_WARPS][G]
  float* s_l = s_m + DC_WARPS * G;                       // [DC_WARPS][G]
  float* c_acc = s_l + DC_WARPS * G;                     // [G][hd]   CTA-level partial (read by peers)
  float* c_m = c_acc + (size_t)G * hd;                   // [G]
  float* c_l = c_m + G;                                  // [G]
  __shared__ int s_pt[DC_MAX_PT];

  const int n_pages = (T_len + page_size - 1) / page_size;
  const bool pt_in_smem = n_pages <= DC_MAX_PT;
  if (pt_in_smem)
    for (int i = threadIdx.x; i < n_pages; i += DC_WARPS * 32) s_pt[i] = page_table[(size_t)b * pt_stride + i];

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
  float m_run = -INFINITY, l_run = 0.f;   // transposed path: lane L keeps the state of head L % G
#pragma unroll
  for (int g = 0; g < G; ++g) {
    m[g] = -INFINITY; l[g] = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
      for (int i = 0; i < V; ++i) acc[g][c][i] = 0.f;
  }
  __syncthreads();

  constexpr int NW = DC_NS * DC_WARPS;
  const int gw = rank * DC_WARPS + wid;
  for (int j0 = gw; j0 < T_len; j0 += NW * TB) {
    uint4 kr[TB][NCH], vr[TB][NCH];
#pragma unroll
    for (int t = 0; t < TB; ++t) {
      const int j = j0 + t * NW;
      if (j < T_len) {
        const int page = pt_in_smem ? s_pt[j / page_size] : page_table[(size_t)b * pt_stride + j / page_size];
        const size_t row = ((size_t)page * page_size + (j % page_size)) * (size_t)row_elems + (size_t)kvh * hd;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const int d = (c * 32 + lane) * V;
          if (d < hd) { kr[t][c] = ldg_cached(k_pool + row + d); vr[t][c] = ldg_cached(v_pool + row + d); }
          else { kr[t][c] = make_uint4(0, 0, 0, 0); vr[t][c] = make_uint4(0, 0, 0, 0); }
        }
      } else {
#pragma unroll
        for (int c = 0; c < NCH; ++c) { kr[t][c] = make_uint4(0, 0, 0, 0); vr[t][c] = make_uint4(0, 0, 0, 0); }
      }
    }
    float s[TB][G];
#pragma unroll
    for (int t = 0; t < TB; ++t) {
      float kf[NCH][V];
#pragma unroll
      for (int c = 0; c < NCH; ++c) unpack<T>(kr[t][c], kf[c]);
#pragma unroll
      for (int g = 0; g < G; ++g) {
        float a = 0.f;
#pragma unroll
        for (int c = 0; c < NCH; ++c)
#pragma unroll
          for (int i = 0; i < V; ++i) a = fmaf(qf[g][c][i], kf[c][i], a);
        s[t][g] = a;
      }
    }
    if constexpr (TB * G == 32) {
      // 32 (token, head) partial dots per lane -> transposed butterfly: 31 shuffles leave lane L with the full
      // sum of pair L = t*G + g (instead of 32 x 5 shuffles with every lane holding every sum), so the scale /
      // max / exp work is done once per pair, not 32 times.
      float sv[32];
#pragma unroll
      for (int t = 0; t < TB; ++t)
#pragma unroll
        for (int g = 0; g < G; ++g) sv[t * G + g] = s[t][g];
#pragma unroll
      for (int w = 16; w >= 1; w >>= 1) {
        const bool up = (lane & w) != 0;
#pragma unroll
        for (int i = 0; i < w; ++i) {
          const float keep = up ? sv[i + w] : sv[i];
          const float send = up ? sv[i] : sv[i + w];
          sv[i] = keep + __shfl_xor_sync(0xffffffffu, send, w);
        }
      }
      const int my_t = lane / G;
      // matmul output rounded to the model dtype, then "/ sqrt(head_dim)" (modeling_gemma.py:266); a power-of-two
      // divisor is applied as an exact multiply (same bits, no division routine)
      const float sr = rnd<T>(sv[0]);
      float sc = (scale_mul != 0.f) ? sr * scale_mul : rnd<T>(sr / scale_div);
      sc = (j0 + my_t * NW < T_len) ? sc : -INFINITY;
      float mx = sc;
#pragma unroll
      for (int o = G; o < 32; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));  // over the TB tokens of my head
      const float mn = fmaxf(m_run, mx);
      const float corr = __expf(m_run - mn);
      const float pr = __expf(sc - mn);                      // 0 for padded tokens
      float psum = pr;
#pragma unroll
      for (int o = G; o < 32; o <<= 1) psum += __shfl_xor_sync(0xffffffffu, psum, o);
      m_run = mn;
      l_run = l_run * corr + psum;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float cg = __shfl_sync(0xffffffffu, corr, g);
#pragma unroll
        for (int c = 0; c < NCH; ++c)
#pragma unroll
          for (int i = 0; i < V; ++i) acc[g][c][i] *= cg;
      }
#pragma unroll
      for (int t = 0; t < TB; ++t) {
        float vf[NCH][V];
#pragma unroll
        for (int c = 0; c < NCH; ++c) unpack<T>(vr[t][c], vf[c]);
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const float pg = __shfl_sync(0xffffffffu, pr, t * G + g);
#pragma unroll
          for (int c = 0; c < NCH; ++c)
#pragma unroll
            for (int i = 0; i < V; ++i) acc[g][c][i] = fmaf(pg, vf[c][i], acc[g][c][i]);
        }
      }
    } else {
#pragma unroll
      for (int t = 0; t < TB; ++t)
  #pragma unroll
        for (int g = 0; g < G; ++g) s[t][g] = warp_sum(s[t][g]);
  #pragma unroll
      for (int g = 0; g < G; ++g) {
        float mn = m[g];
  #pragma unroll
        for (int t = 0; t < TB; ++t) {
          // matmul output rounded to the model dtype, then "/ sqrt(head_dim)" (modeling_gemma.py:266)
          // (a power-of-two divisor is applied as an exact multiply: same bits, no division routine)
          const float sr = rnd<T>(s[t][g]);
          const float sc = (scale_mul != 0.f) ? sr * scale_mul : rnd<T>(sr / scale_div);
          s[t][g] = (j0 + t * NW < T_len) ? sc : -INFINITY;
          mn = fmaxf(mn, s[t][g]);
        }
        const float corr = __expf(m[g] - mn);
        m[g] = mn;
        l[g] *= corr;
  #pragma unroll
        for (int c = 0; c < NCH; ++c)
  #pragma unroll
          for (int i = 0; i < V; ++i) acc[g][c][i] *= corr;
  #pragma unroll
        for (int t = 0; t < TB; ++t) s[t][g] = __expf(s[t][g] - mn);  // p; 0 for the padded tokens
  #pragma unroll
        for (int t = 0; t < TB; ++t) l[g] += s[t][g];
      }
  #pragma unroll
      for (int t = 0; t < TB; ++t) {
        float vf[NCH][V];
  #pragma unroll
        for (int c = 0; c < NCH; ++c) unpack<T>(vr[t][c], vf[c]);
  #pragma unroll
        for (int g = 0; g < G; ++g)
  #pragma unroll
          for (int c = 0; c < NCH; ++c)
  #pragma unroll
            for (int i = 0; i < V; ++i) acc[g][c][i] = fmaf(s[t][g], vf[c][i], acc[g][c][i]);
      }
    }
  }
  if constexpr (TB * G == 32) {  // the per-head running (m, l) live in the lanes of that head: gather them
#pragma unroll
    for (int g = 0; g < G; ++g) { m[g] = __shfl_sync(0xffffffffu, m_run, g); l[g] = __shfl_sync(0xffffffffu, l_run, g); }
  }

  // ---- merge the warps of this CTA
#pragma unroll
  for (int g = 0; g < G; ++g) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int d = (c * 32 + lane) * V;
      if (d < hd)
#pragma unroll
        for (int i = 0; i < V; i += 4)
          *reinterpret_cast<float4*>(&s_acc[((size_t)wid * G + g) * hd + d + i]) =
              make_float4(acc[g][c][i], acc[g][c][i + 1], acc[g][c][i + 2], acc[g][c][i + 3]);
    }
    if (lane == 0) { s_m[wid * G + g] = m[g]; s_l[wid * G + g] = l[g]; }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < G * hd; e += DC_WARPS * 32) {
    const int g = e / hd, d = e % hd;
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < DC_WARPS; ++w) M = fmaxf(M, s_m[w * G + g]);
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < DC_WARPS; ++w) {
      const float mw = s_m[w * G + g];
      if (mw != -INFINITY) a = fmaf(__expf(mw - M), s_acc[((size_t)w * G + g) * hd + d], a);
    }
    c_acc[e] = a;
  }
  if (threadIdx.x < G) {
    const int g = threadIdx.x;
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < DC_WARPS; ++w) M = fmaxf(M, s_m[w * G + g]);
    float L = 0.f;
#pragma unroll
    for (int w = 0; w < DC_WARPS; ++w) {
      const float mw = s_m[w * G + g];
      if (mw != -INFINITY) L = fmaf(__expf(mw - M), s_l[w * G + g], L);
    }
    c_m[g] = M;
    c_l[g] = L;
  }
  cluster.sync();  // every CTA's partial is visible cluster-wide

  // ---- merge the CTAs: CTA `rank` finalises its slice of the G*hd outputs through DSMEM
  const int per = (G * hd + DC_NS - 1) / DC_NS;
  const int e_end = min(G * hd, (rank + 1) * per);
  for (int e = rank * per + threadIdx.x; e < e_end; e += DC_WARPS * 32) {
    const int g = e / hd;
    float pm[DC_NS], pl[DC_NS], pa[DC_NS];
#pragma unroll
    for (int r = 0; r < DC_NS; ++r) {
      pm[r] = *cluster.map_shared_rank(c_m + g, r);
      pl[r] = *cluster.map_shared_rank(c_l + g, r);
      pa[r] = *cluster.map_shared_rank(c_acc + e, r);
    }
    float M = -INFINITY;
#pragma unroll
    for (int r = 0; r < DC_NS; ++r) M = fmaxf(M, pm[r]);
    float a = 0.f, L = 0.f;
#pragma unroll
    for (int r = 0; r < DC_NS; ++r)
      if (pm[r] != -INFINITY) {
        const float w = __expf(pm[r] - M);
        a = fmaf(w, pa[r], a);
        L = fmaf(w, pl[r], L);
      }
    out[(size_t)b * nq * hd + (size_t)(kvh * G + g) * hd + (e % hd)] = from_f<T>(a / L);
  }
  cluster.sync();  // nobody leaves while a peer may still read its shared memory
}

template <typename T, int G, int NCH, int DC_NS>
static int launch_da_ns(void* out, const void* q, const void* k_pool, const void* v_pool, const int32_t* page_table,
                     int pt_stride, int page_size, const int32_t* kv_len, int kv_len_add, int B, int nq, int nkv,
                     int hd, float scale_div, Prefetch pf, cudaStream_t st) {
  constexpr int TB = NCH == 1 ? 4 : 2;
  auto kern = decode_attention_cluster_kernel<T, G, NCH, TB, DC_NS>;
  if (DC_NS > 8) cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  // x / 2^k == x * 2^-k exactly (no underflow at these magnitudes): skip the IEEE division routine
  int ex = 0;
  const float scale_mul = (frexpf(scale_div, &ex) == 0.5f) ? 1.0f / scale_div : 0.f;
  const size_t smem = ((size_t)DC_WARPS * G * hd + 2 * DC_WARPS * G + (size_t)G * hd + 2 * G) * sizeof(float);
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    set_error("decode_attention: cannot reserve %zu B of shared memory", smem);
    cudaGetLastError();
    return PG_ERR_CUDA;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(DC_NS, B, nkv);
  cfg.blockDim = dim3(DC_WARPS * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  static const int pdl = env_int("PG_PDL", 1);
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = DC_NS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, (T*)out, (const T*)q, (const T*)k_pool, (const T*)v_pool,
                                     page_table, pt_stride, page_size, kv_len, kv_len_add, nq, nkv, hd, scale_div,
                                     scale_mul, pf);
  if (e != cudaSuccess) {
    set_error("decode_attention launch: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return PG_ERR_CUDA;
  }
  return check_launch("decode_attention");
}

template <typename T, int G, int NCH>
static int launch_da(void* out, const void* q, const void* k_pool, const void* v_pool, const int32_t* page_table,
                     int pt_stride, int page_size, const int32_t* kv_len, int kv_len_add, int B, int nq, int nkv,
                     int hd, float scale_div, Prefetch pf, cudaStream_t st) {
  // Cluster size: as many CTAs per (sequence, kv head) as keeps the whole launch within one wave of the 148 SMs
  // (the kernel needs a full SM: 255 registers x 256 threads).  16 is a non-portable cluster size.
  static const int ns_env = env_int("PG_ATTN_NS", 0);
  int ns = ns_env;
  if (!ns) {
    const int seqs = B * nkv;
    ns = seqs <= 9 ? 16 : (seqs <= 18 ? 8 : (seqs <= 37 ? 4 : (seqs <= 74 ? 2 : 1)));
  }
#define PG_DA_NS(N_) \
  return launch_da_ns<T, G, NCH, N_>(out, q, k_pool, v_pool, page_table, pt_stride, page_size, kv_len, kv_len_add, B, nq, nkv, hd, scale_div, pf, st)
  switch (ns) {
    case 16: PG_DA_NS(16);
    case 8: PG_DA_NS(8);
    case 4: PG_DA_NS(4);
    case 2: PG_DA_NS(2);
    default: PG_DA_NS(1);
  }
#undef PG_DA_NS
}

}  // namespace pg

using namespace pg;

extern "C" {

long long pg_decode_attention_ws_floats(int, int, int, int) { return 0; }

int pg_decode_attention(void* out, const void* q, const void* k_pool, const void* v_pool,
                        const int32_t* page_table, int pt_stride, int page_size, const int32_t* kv_len,
                        int kv_len_add, int B, int nq, int nkv, int hd, float scale_div, float* /*ws*/,
                        int* /*counters*/, int /*max_splits*/, int dtype, void* stream) {
  const Prefetch pf = take_prefetch();
  PG_REQUIRE(B > 0 && B <= 65535 && nq % nkv == 0, "decode_attention: bad shape");
  const int G = nq / nkv;
  cudaStream_t st = (cudaStream_t)stream;
#define PG_DA(GG, NCH) \
  return launch_da<T, GG, NCH>(out, q, k_pool, v_pool, page_table, pt_stride, page_size, kv_len, kv_len_add, B, nq, nkv, hd, scale_div, pf, st)
  PG_DISPATCH_DTYPE(dtype, T, {
    constexpr int V = Vec<T>::N;
    PG_REQUIRE(hd % V == 0 && hd % 4 == 0 && hd <= 64 * V, "decode_attention: unsupported head_dim %d", hd);
    const int nch = (hd + 32 * V - 1) / (32 * V);
    if (nch == 1) {
      switch (G) {
        case 1: PG_DA(1, 1);
        case 2: PG_DA(2, 1);
        case 4: PG_DA(4, 1);
        case 8: PG_DA(8, 1);
        default: set_error("decode_attention: unsupported group size %d", G); return PG_ERR_INVALID;
      }
    } else {
      switch (G) {
        case 1: PG_DA(1, 2);
        case 2: PG_DA(2, 2);
        case 4: PG_DA(4, 2);
        case 8: PG_DA(8, 2);
        default: set_error("decode_attention: unsupported group size %d", G); return PG_ERR_INVALID;
      }
    }
  });
#undef PG_DA
  return PG_OK;
}

}  // extern "C"
