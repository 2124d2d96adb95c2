// One-shot all-reduce over NVLink peer memory for the tensor-parallel decoder's small messages
// ((B, 2048) after o_proj and down_proj: 4 KB at batch 1, 36 per token).  NCCL's ring/tree latency
// (~14 us per call on B200/NVSwitch at this size) is most of a tensor-parallel decode step; here every
// rank publishes its partial in a symmetric (peer-mapped) buffer, raises a flag in every peer's memory,
// waits for the peers' flags and sums all partials itself, in rank order, so every rank gets bit-identical
// results.  One launch per all-reduce per GPU; ranks run on different GPUs (never two waiting kernels on
// one device).  Double-buffered by step parity; flags carry a monotonically increasing step number, so
// nothing is ever reset and a captured CUDA graph can be replayed indefinitely.  A bounded spin turns a
// lost peer into an error flag instead of a hang.
#include "common.cuh"

namespace pg {

constexpr int AR_THREADS = 512;
constexpr int AR_MAX_RANKS = 16;

__device__ __forceinline__ uint4 ld_volatile_v4(const void* p) {
  uint4 r;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}

template <typename T>
__global__ void __launch_bounds__(AR_THREADS)
allreduce_oneshot_kernel(T* __restrict__ x, char* const* __restrict__ peers, int rank, int tp, int n, long long cap,
                         int* __restrict__ step_counter, int* __restrict__ err) {
  constexpr int V = Vec<T>::N;
  __shared__ int s_step;
  if (threadIdx.x == 0) s_step = *step_counter;
  __syncthreads();
  const int step = s_step, slot = step & 1;
  const unsigned target = (unsigned)step + 1u;
  char* mine = peers[rank];
  T* my_data = reinterpret_cast<T*>(mine + (long long)slot * cap);
  // 1. publish this rank's partial
  for (int i = threadIdx.x * V; i < n; i += AR_THREADS * V)
    *reinterpret_cast<uint4*>(my_data + i) = *reinterpret_cast<const uint4*>(x + i);
  __threadfence_system();
  __syncthreads();
  // 2. tell every peer (and ourselves) that slot `slot` holds step `step`
  if (threadIdx.x < tp) {
    unsigned* f = reinterpret_cast<unsigned*>(peers[threadIdx.x] + 2 * cap) + slot * AR_MAX_RANKS + rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(target) : "memory");
  }
  // 3. wait until every rank's partial for this step is published
  if (threadIdx.x < tp) {
    const unsigned* f = reinterpret_cast<const unsigned*>(mine + 2 * cap) + slot * AR_MAX_RANKS + threadIdx.x;
    unsigned v = 0;
    long long spins = 0;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
    } while ((int)(v - target) < 0 && ++spins < (1LL << 27));
    if ((int)(v - target) < 0 && err) *err = 2;  // a peer never arrived: flag it instead of hanging
  }
  __syncthreads();
  // 4. sum the partials in rank order (identical on every rank)
  for (int i = threadIdx.x * V; i < n; i += AR_THREADS * V) {
    float acc[V];
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] = 0.f;
    for (int r = 0; r < tp; ++r) {
      float f[V];
      unpack<T>(ld_volatile_v4(reinterpret_cast<const T*>(peers[r] + (long long)slot * cap) + i), f);
#pragma unroll
      for (int k = 0; k < V; ++k) acc[k] += f[k];
    }
    *reinterpret_cast<uint4*>(x + i) = pack<T>(acc);
  }
  if (threadIdx.x == 0) *step_counter = (int)target;
}

}  // namespace pg

using namespace pg;

extern "C" int pg_allreduce_oneshot(void* x, const void* peers_dev, int rank, int tp, int n, long long cap_bytes,
                                    int* step_counter, int* err_flag, int dtype, void* stream) {
  PG_REQUIRE(tp >= 2 && tp <= AR_MAX_RANKS && rank >= 0 && rank < tp, "allreduce_oneshot: bad rank/size %d/%d", rank, tp);
  PG_DISPATCH_DTYPE(dtype, T, {
    PG_REQUIRE(n % Vec<T>::N == 0 && (long long)n * (long long)sizeof(T) <= cap_bytes && cap_bytes % 16 == 0,
               "allreduce_oneshot: n=%d does not fit the symmetric buffer", n);
    allreduce_oneshot_kernel<T><<<1, AR_THREADS, 0, (cudaStream_t)stream>>>((T*)x, (char* const*)peers_dev, rank, tp, n, cap_bytes,
                                                                            step_counter, err_flag);
  });
  return check_launch("allreduce_oneshot");
}
