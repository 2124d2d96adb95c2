// softmax(logits / temperature) + nucleus sampling without a full sort.
// Reference: inference.py:15-24,65-66 — sort descending, keep while (cumsum - p_i) <= top_p,
// renormalise, multinomial.  An element is kept iff the mass of the strictly larger
// probabilities is <= top_p (ties: lower index first, as a stable descending sort orders them),
// so the nucleus is found by a 4-level radix select over the fp32 bit patterns with per-bin
// probability mass, then one index-order scan draws the sample.  One CTA per batch row; the
// [V] fp32 row (1 MB at V=257216) stays L2 resident across the passes.
#include "common.cuh"

namespace pg {

constexpr int TP_THREADS = 1024;

struct Philox {
  // Philox4x32-10, counter-based: (seed, subsequence) -> 4 x u32
  static __device__ __forceinline__ void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
  static __device__ float uniform(unsigned long long seed, unsigned long long subseq) {
    uint32_t c[4] = {(uint32_t)subseq, (uint32_t)(subseq >> 32), 0u, 0u};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int i = 0; i < 10; ++i) { round(c, k0, k1); k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
    return (float)(c[0] >> 8) * (1.0f / 16777216.0f);  // [0,1)
  }
};

__device__ __forceinline__ float block_reduce_max(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = (lane < (TP_THREADS / 32)) ? red[lane] : -INFINITY;
  return warp_max(t);
}

// exclusive block scan of one float per thread; returns the exclusive prefix, *total = sum
__device__ __forceinline__ float block_exclusive_scan(float v, float* red, float* total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();
  if (lane == 31) red[w] = inc;
  __syncthreads();
  if (w == 0) {
    float x = red[lane];
    float xi = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      float t = __shfl_up_sync(0xffffffffu, xi, o);
      if (lane >= o) xi += t;
    }
    red[lane] = xi - x;          // exclusive warp offsets
    if (lane == 31) red[32] = xi;  // total
  }
  __syncthreads();
  *total = red[32];
  return red[w] + (inc - v);
}

__global__ void __launch_bounds__(TP_THREADS)
top_p_kernel(int64_t* __restrict__ out, const float* __restrict__ logits, float* __restrict__ probs_ws,
             long long V, float temperature, float top_p, unsigned long long seed,
             const int* __restrict__ rng_offset, int* __restrict__ nucleus_size) {
  __shared__ float red[33];
  __shared__ float s_mass[256];
  __shared__ int s_cnt[256];
  __shared__ uint32_t s_prefix;
  __shared__ float s_R;
  __shared__ int s_found, s_tiecnt;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  const float* lrow = logits + (size_t)b * V;
  float* prow = probs_ws + (size_t)b * V;

  // ---- softmax(logits / temperature), fp32 (inference.py:65)
  float mx = -INFINITY;
  for (long long i = tid; i < V; i += TP_THREADS) mx = fmaxf(mx, lrow[i] / temperature);
  mx = block_reduce_max(mx, red);
  float sum = 0.f;
  for (long long i = tid; i < V; i += TP_THREADS) {
    const float e = expf(lrow[i] / temperature - mx);
    prow[i] = e;
    sum += e;
  }
  sum = block_sum(sum, red);
  for (long long i = tid; i < V; i += TP_THREADS) prow[i] = prow[i] / sum;
  __syncthreads();

  // ---- radix select of the boundary value
  if (tid == 0) { s_prefix = 0u; s_R = 0.f; }
  bool keep_all = false;
  for (int level = 0; level < 4; ++level) {
    const int shift = 24 - 8 * level;
    for (int i = tid; i < 256; i += TP_THREADS) { s_mass[i] = 0.f; s_cnt[i] = 0; }
    __syncthreads();
    const uint32_t prefix = s_prefix;
    const long long Vr = ((V + 31) / 32) * 32;  // keep whole warps in the loop for the shuffles
    for (long long i = tid; i < Vr; i += TP_THREADS) {
      float p = 0.f;
      int bin = -1;
      if (i < V) {
        p = prow[i];
        const uint32_t bits = __float_as_uint(p);
        if (level == 0 || (bits >> (shift + 8)) == prefix) bin = (int)((bits >> shift) & 255u);
      }
      // warp-aggregated accumulation: one shared-memory atomic per distinct bin per warp
      unsigned todo = __ballot_sync(0xffffffffu, bin >= 0);
      while (todo) {
        const int leader = __ffs(todo) - 1;
        const int lb = __shfl_sync(0xffffffffu, bin, leader);
        const bool mine = (bin == lb);
        const unsigned grp = __ballot_sync(0xffffffffu, mine);
        const float m = warp_sum(mine ? p : 0.f);
        if (lane == leader) { atomicAdd(&s_mass[lb], m); atomicAdd(&s_cnt[lb], __popc(grp)); }
        todo &= ~grp;
      }
    }
    __syncthreads();
    if (tid == 0) {
      float R = s_R;
      int found = -1;
      for (int bin = 255; bin >= 0; --bin) {
        if (s_cnt[bin] == 0) continue;
        if (R + s_mass[bin] > top_p) { found = bin; break; }
        R += s_mass[bin];
      }
      s_found = found;
      if (found >= 0) {
        s_R = R;
        s_prefix = (prefix << 8) | (uint32_t)found;
        s_tiecnt = s_cnt[found];
      }
    }
    __syncthreads();
    if (s_found < 0) {
      // level 0: the whole distribution fits under top_p.  Deeper levels: fp32 summation-order
      // noise hid the boundary; everything inside the current prefix is kept.
      keep_all = true;
      if (tid == 0) s_prefix = (level == 0) ? 0u : (prefix << (8 * (4 - level)));
      __syncthreads();
      break;
    }
  }
  const uint32_t vbits = s_prefix;
  const float vstar = __uint_as_float(vbits);
  int n_tie_keep;
  if (keep_all) n_tie_keep = 0x7fffffff;
  else {
    const float room = top_p - s_R;
    long long n = (room < 0.f) ? 0 : (long long)floorf(room / vstar) + 1;
    n_tie_keep = (int)(n > s_tiecnt ? s_tiecnt : n);
    if (n_tie_keep < 1) n_tie_keep = 1;
  }

  // ---- index-order scan: tie ranks, kept mass, draw
  const long long seg = (V + TP_THREADS - 1) / TP_THREADS;
  const long long i0 = (long long)tid * seg, i1 = (i0 + seg < V) ? i0 + seg : V;
  float ties = 0.f;
  for (long long i = i0; i < i1; ++i) ties += (__float_as_uint(prow[i]) == vbits) ? 1.f : 0.f;
  float tie_total;
  float tie_before = block_exclusive_scan(ties, red, &tie_total);
  float mass = 0.f, cntf = 0.f;
  {
    int rank = (int)(tie_before + 0.5f);
    for (long long i = i0; i < i1; ++i) {
      const uint32_t bits = __float_as_uint(prow[i]);
      bool keep = bits > vbits;
      if (bits == vbits) { keep = keep_all || (rank < n_tie_keep); ++rank; }
      if (keep) { mass += prow[i]; cntf += 1.f; }
    }
  }
  float Z;
  const float mass_before = block_exclusive_scan(mass, red, &Z);
  float kept_total;
  block_exclusive_scan(cntf, red, &kept_total);
  if (tid == 0 && nucleus_size) nucleus_size[b] = (int)(kept_total + 0.5f);
  const float u = Philox::uniform(seed, (unsigned long long)(rng_offset ? *rng_offset : 0) * 4096ull + b);
  const float target = u * Z;
  // the owner is the last thread whose exclusive prefix is <= target and that holds kept mass
  __shared__ int s_owner;
  if (tid == 0) s_owner = -1;
  __syncthreads();
  if (mass > 0.f && mass_before <= target) atomicMax(&s_owner, tid);
  __syncthreads();
  if (tid == s_owner) {
    int rank = (int)(tie_before + 0.5f);
    float run = mass_before;
    long long pick = -1, last_kept = -1;
    for (long long i = i0; i < i1; ++i) {
      const uint32_t bits = __float_as_uint(prow[i]);
      bool keep = bits > vbits;
      if (bits == vbits) { keep = keep_all || (rank < n_tie_keep); ++rank; }
      if (keep) {
        last_kept = i;
        run += prow[i];
        if (run > target) { pick = i; break; }
      }
    }
    out[b] = (pick >= 0) ? pick : last_kept;
  }
}


// ------------------------------------------------------------------------------------------------
// Cluster version: TPC_NC CTAs per batch row, each keeps its slice of the row in shared memory for every pass
// (softmax, 4 radix levels, index-order scan), so the row is read from global exactly once and 8x more SMs
// work on it; row-wide quantities (max, sum, 256-bin mass/count histograms, tie counts, kept mass) meet
// through distributed shared memory, always summed in rank order so every CTA takes identical decisions.
constexpr int TPC_NC = 8, TPC_THREADS = 512, TPC_WARPS = TPC_THREADS / 32;

struct TpcPub {                    // what a CTA publishes to its cluster peers
  float mass[256];
  int cnt[256];
  float scalar[4];                 // [0] max, [1] sum, [2] local tie count, [3] local kept mass
  float kept_cnt;
};

__device__ __forceinline__ float tpc_block_max(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = (lane < TPC_WARPS) ? red[lane] : -INFINITY;
  return warp_max(t);
}
__device__ __forceinline__ float tpc_block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = (lane < TPC_WARPS) ? red[lane] : 0.f;
  return warp_sum(t);
}
// exclusive scan of one float per thread over TPC_THREADS threads; *total = block sum
__device__ __forceinline__ float tpc_block_exclusive_scan(float v, float* red, float* total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();
  if (lane == 31) red[w] = inc;
  __syncthreads();
  if (w == 0) {
    float x = (lane < TPC_WARPS) ? red[lane] : 0.f;
    float xi = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      float t = __shfl_up_sync(0xffffffffu, xi, o);
      if (lane >= o) xi += t;
    }
    if (lane < TPC_WARPS) red[lane] = xi - x;
    if (lane == 31) red[32] = xi;
  }
  __syncthreads();
  *total = red[32];
  return red[w] + (inc - v);
}
__device__ __forceinline__ void tpc_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <typename V>
__device__ __forceinline__ V tpc_ld_peer(const V* local, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"((uint32_t)__cvta_generic_to_shared(local)), "r"(rank));
  uint32_t bits;
  asm volatile("ld.shared::cluster.b32 %0, [%1];" : "=r"(bits) : "r"(remote) : "memory");
  return *reinterpret_cast<V*>(&bits);
}

__global__ void __launch_bounds__(TPC_THREADS, 1)
top_p_cluster_kernel(int64_t* __restrict__ out, const float* __restrict__ logits, float* __restrict__ probs_ws,
                     long long V, int chunk, float temperature, float top_p, unsigned long long seed,
                     const int* __restrict__ rng_offset, int* __restrict__ nucleus_size) {
  extern __shared__ __align__(16) float s_p[];          // this CTA's slice of the row
  __shared__ TpcPub pub;
  __shared__ float red[33];
  __shared__ float s_mass[256];
  __shared__ int s_cnt[256];
  __shared__ float w_mass[TPC_WARPS][256];               // per-warp private histograms (no cross-warp contention)
  __shared__ int w_cnt[TPC_WARPS][256];
  __shared__ uint32_t s_prefix;
  __shared__ float s_R;
  __shared__ int s_found, s_tiecnt, s_owner;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int b = blockIdx.y, tid = threadIdx.x, wid = tid >> 5;
  const long long g0 = (long long)rank * chunk;
  const int n_local = (int)max(0ll, min((long long)chunk, V - g0));
  const float* lrow = logits + (size_t)b * V + g0;
  float* prow = probs_ws + (size_t)b * V + g0;

  // ---- softmax(logits / temperature), fp32 (inference.py:65)
  float mx = -INFINITY;
  for (int i = tid; i < n_local; i += TPC_THREADS) {
    const float x = lrow[i] / temperature;
    s_p[i] = x;
    mx = fmaxf(mx, x);
  }
  mx = tpc_block_max(mx, red);
  if (tid == 0) pub.scalar[0] = mx;
  tpc_cluster_sync();
  mx = -INFINITY;
#pragma unroll
  for (int r = 0; r < TPC_NC; ++r) mx = fmaxf(mx, tpc_ld_peer(&pub.scalar[0], r));
  float sum = 0.f;
  for (int i = tid; i < n_local; i += TPC_THREADS) {
    const float e = expf(s_p[i] - mx);
    s_p[i] = e;
    sum += e;
  }
  sum = tpc_block_sum(sum, red);
  if (tid == 0) pub.scalar[1] = sum;
  tpc_cluster_sync();
  sum = 0.f;
#pragma unroll
  for (int r = 0; r < TPC_NC; ++r) sum += tpc_ld_peer(&pub.scalar[1], r);
  for (int i = tid; i < n_local; i += TPC_THREADS) {
    const float pr = s_p[i] / sum;
    s_p[i] = pr;
    prow[i] = pr;
  }
  if (tid == 0) { s_prefix = 0u; s_R = 0.f; }
  __syncthreads();

  // ---- radix select of the boundary value (4 levels of 8 bits over the fp32 bit patterns)
  bool keep_all = false;
  for (int level = 0; level < 4; ++level) {
    const int shift = 24 - 8 * level;
    for (int i = tid; i < TPC_WARPS * 256; i += TPC_THREADS) { (&w_mass[0][0])[i] = 0.f; (&w_cnt[0][0])[i] = 0; }
    __syncthreads();
    const uint32_t prefix = s_prefix;
    for (int i = tid; i < n_local; i += TPC_THREADS) {
      const float pr = s_p[i];
      const uint32_t bits = __float_as_uint(pr);
      if (level == 0 || (bits >> (shift + 8)) == prefix) {
        const int bin = (int)((bits >> shift) & 255u);
        atomicAdd(&w_mass[wid][bin], pr);
        atomicAdd(&w_cnt[wid][bin], 1);
      }
    }
    __syncthreads();
    if (tid < 256) {
      float m = 0.f;
      int c = 0;
#pragma unroll
      for (int w = 0; w < TPC_WARPS; ++w) { m += w_mass[w][tid]; c += w_cnt[w][tid]; }
      pub.mass[tid] = m;
      pub.cnt[tid] = c;
    }
    tpc_cluster_sync();
    if (tid < 256) {
      float m = 0.f;
      int c = 0;
#pragma unroll
      for (int r = 0; r < TPC_NC; ++r) { m += tpc_ld_peer(&pub.mass[tid], r); c += tpc_ld_peer(&pub.cnt[tid], r); }
      s_mass[tid] = m;
      s_cnt[tid] = c;
    }
    __syncthreads();
    if (tid == 0) {
      float R = s_R;
      int found = -1;
      for (int bin = 255; bin >= 0; --bin) {
        if (s_cnt[bin] == 0) continue;
        if (R + s_mass[bin] > top_p) { found = bin; break; }
        R += s_mass[bin];
      }
      s_found = found;
      if (found >= 0) {
        s_R = R;
        s_prefix = (prefix << 8) | (uint32_t)found;
        s_tiecnt = s_cnt[found];
      }
    }
    tpc_cluster_sync();   // peers have read this CTA's histogram; also orders the shared-memory decision
    if (s_found < 0) {
      keep_all = true;
      if (tid == 0) s_prefix = (level == 0) ? 0u : (prefix << (8 * (4 - level)));
      __syncthreads();
      break;
    }
  }
  const uint32_t vbits = s_prefix;
  const float vstar = __uint_as_float(vbits);
  int n_tie_keep;
  if (keep_all) n_tie_keep = 0x7fffffff;
  else {
    const float room = top_p - s_R;
    long long n = (room < 0.f) ? 0 : (long long)floorf(room / vstar) + 1;
    n_tie_keep = (int)(n > s_tiecnt ? s_tiecnt : n);
    if (n_tie_keep < 1) n_tie_keep = 1;
  }

  // ---- index-order scan: tie ranks, kept mass, draw.  Thread t owns the contiguous segment [t*seg, (t+1)*seg) of the
  // slice; seg is odd, so the 32 lanes of a warp walk 32 different banks.
  const int seg = ((chunk + TPC_THREADS - 1) / TPC_THREADS) | 1;
  const int i0 = min(tid * seg, n_local), i1 = min(i0 + seg, n_local);
  float ties = 0.f;
  for (int i = i0; i < i1; ++i) ties += (__float_as_uint(s_p[i]) == vbits) ? 1.f : 0.f;
  float tie_local_total;
  const float tie_before_local = tpc_block_exclusive_scan(ties, red, &tie_local_total);
  if (tid == 0) pub.scalar[2] = tie_local_total;
  tpc_cluster_sync();
  float tie_before_cta = 0.f;
  for (uint32_t r = 0; r < rank; ++r) tie_before_cta += tpc_ld_peer(&pub.scalar[2], r);
  const float tie_before = tie_before_cta + tie_before_local;
  float mass = 0.f, cntf = 0.f;
  {
    int trank = (int)(tie_before + 0.5f);
    for (int i = i0; i < i1; ++i) {
      const uint32_t bits = __float_as_uint(s_p[i]);
      bool keep = bits > vbits;
      if (bits == vbits) { keep = keep_all || (trank < n_tie_keep); ++trank; }
      if (keep) { mass += s_p[i]; cntf += 1.f; }
    }
  }
  float mass_local_total, cnt_local_total;
  const float mass_before_local = tpc_block_exclusive_scan(mass, red, &mass_local_total);
  tpc_block_exclusive_scan(cntf, red, &cnt_local_total);
  if (tid == 0) { pub.scalar[3] = mass_local_total; pub.kept_cnt = cnt_local_total; }
  tpc_cluster_sync();
  float cta_mass[TPC_NC], Z = 0.f, kept_total = 0.f, mass_before_cta = 0.f;
#pragma unroll
  for (int r = 0; r < TPC_NC; ++r) {
    cta_mass[r] = tpc_ld_peer(&pub.scalar[3], r);
    if ((uint32_t)r == rank) mass_before_cta = Z;
    Z += cta_mass[r];
    kept_total += tpc_ld_peer(&pub.kept_cnt, r);
  }
  if (rank == 0 && tid == 0 && nucleus_size) nucleus_size[b] = (int)(kept_total + 0.5f);
  const float u = Philox::uniform(seed, (unsigned long long)(rng_offset ? *rng_offset : 0) * 4096ull + b);
  const float target = u * Z;
  // the owner is the last (CTA, thread) whose exclusive prefix is <= target and that holds kept mass
  bool cta_owns = cta_mass[rank] > 0.f && mass_before_cta <= target;
  {
    float pre = 0.f;
#pragma unroll
    for (int r = 0; r < TPC_NC; ++r) {
      if ((uint32_t)r > rank && cta_mass[r] > 0.f && pre <= target) cta_owns = false;
      pre += cta_mass[r];
    }
  }
  if (tid == 0) s_owner = -1;
  __syncthreads();
  const float mass_before = mass_before_cta + mass_before_local;
  if (cta_owns && mass > 0.f && mass_before <= target) atomicMax(&s_owner, tid);
  __syncthreads();
  if (cta_owns && tid == s_owner) {
    int trank = (int)(tie_before + 0.5f);
    float run = mass_before;
    long long pick = -1, last_kept = -1;
    for (int i = i0; i < i1; ++i) {
      const uint32_t bits = __float_as_uint(s_p[i]);
      bool keep = bits > vbits;
      if (bits == vbits) { keep = keep_all || (trank < n_tie_keep); ++trank; }
      if (keep) {
        last_kept = g0 + i;
        run += s_p[i];
        if (run > target) { pick = g0 + i; break; }
      }
    }
    out[b] = (pick >= 0) ? pick : last_kept;
  }
  tpc_cluster_sync();   // nobody leaves while a peer may still read its published values
}

}  // namespace pg

using namespace pg;

extern "C" int pg_top_p_sample(int64_t* out, const float* logits, float* probs_ws, int B, int64_t V,
                               float temperature, float top_p, unsigned long long seed, const int* rng_offset,
                               int* nucleus_size, void* stream) {
  PG_REQUIRE(B > 0 && V > 0 && temperature > 0.f, "top_p_sample: bad arguments");
  // cluster version: 8 CTAs per row, the row slice stays in shared memory (PG_TOPP_CLUSTER=0: one CTA per row)
  static const int use_cluster = env_int("PG_TOPP_CLUSTER", 1);
  const int chunk = (int)(((V + TPC_NC - 1) / TPC_NC + 3) / 4 * 4);
  const size_t smem = (size_t)chunk * sizeof(float);
  if (use_cluster && B <= 65535 && smem <= 160 * 1024) {
    if (smem > 16 * 1024 &&
        cudaFuncSetAttribute(top_p_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      set_error("top_p_sample: cannot reserve %zu B of shared memory", smem);
      cudaGetLastError();
      return PG_ERR_CUDA;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(TPC_NC, B);
    cfg.blockDim = dim3(TPC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = TPC_NC;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, top_p_cluster_kernel, out, logits, probs_ws, (long long)V, chunk, temperature, top_p,
                                       seed, rng_offset, nucleus_size);
    if (e != cudaSuccess) {
      set_error("top_p_sample launch: %s", cudaGetErrorString(e));
      cudaGetLastError();
      return PG_ERR_CUDA;
    }
    return check_launch("top_p_sample");
  }
  top_p_kernel<<<B, TP_THREADS, 0, (cudaStream_t)stream>>>(out, logits, probs_ws, (long long)V, temperature,
                                                            top_p, seed, rng_offset, nucleus_size);
  return check_launch("top_p_sample");
}
