// Stand-alone pieces of the tensor-parallel exchange (protocol: tp_exchange.cuh): the per-step epoch bump, the
// producer for partials that a tensor-core GEMM left in local memory (batched decode step), the row-wise consumer
// (sum of partials + residual + RMSNorm) for that step, and the (max, index) key exchange of the vocabulary-split
// lm_head.  The batch <= 8 GEMV step has its producer / consumer fused into the GEMV kernels (gemv.cu).
#include "tp_exchange.cuh"

namespace pg {

__global__ void tp_begin_step_kernel(int* epoch) {
  if (threadIdx.x == 0) *epoch = *epoch + 1;
}

// partial[n] (fp32, local) -> word i of this rank's slot in EVERY rank's buffer; consecutive threads write
// consecutive 8-byte words, so a warp's stores to one peer are one 256-byte NVLink write
__global__ void __launch_bounds__(256) tp_push_kernel(const float* __restrict__ partial, long long n, TpEx ex) {
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t seq = tp_seq(ex);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = partial[i];
    for (int p = 0; p < ex.tp; ++p) tp_store_word(tp_slot(ex, p, seq, ex.rank), i, v, seq);
  }
}

// One CTA per row: x_new = rnd(x_in + rnd(sum over ranks of the partials)); out = RMSNorm(x_new) * (1 + w).
// Replaces the all-reduce + residual add + GemmaRMSNorm (modeling_gemma.py:114-120, 318-336) of the batched step.
template <typename T>
__global__ void __launch_bounds__(256)
rmsnorm_reduce_kernel(T* __restrict__ out, T* __restrict__ x_out, const T* __restrict__ x_in, const T* __restrict__ w,
                      int D, float eps, TpEx ex) {
  extern __shared__ __align__(16) float xs[];
  __shared__ float red[32];
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t seq = tp_seq(ex);
  bool dead = tp_failed(ex);
  const size_t row = blockIdx.x;
  float ss = 0.f;
  for (int c2 = threadIdx.x; c2 < D / 2; c2 += blockDim.x) {
    const size_t i = row * D + 2 * c2;
    const float2 p = tp_reduce_pair(ex, seq, (long long)(i >> 1), dead);
    const float v0 = rnd<T>(to_f<T>(x_in[i]) + rnd<T>(p.x));
    const float v1 = rnd<T>(to_f<T>(x_in[i + 1]) + rnd<T>(p.y));
    xs[2 * c2] = v0;
    xs[2 * c2 + 1] = v1;
    x_out[i] = from_f<T>(v0);
    x_out[i + 1] = from_f<T>(v1);
    ss += v0 * v0 + v1 * v1;
  }
  const float tot = block_sum(ss, red);
  const float inv = rsqrtf(tot / (float)D + eps);
  for (int c = threadIdx.x; c < D; c += blockDim.x)
    out[row * D + c] = from_f<T>((xs[c] * inv) * (1.0f + to_f<T>(w[c])));
}

// lm_head is split by vocabulary: every rank holds the packed (value, LOCAL index) argmax key of its shard.
// Re-base the index to the full vocabulary and hand the key to every rank (two 8-byte words per row).
__global__ void tp_keys_push_kernel(const unsigned long long* __restrict__ keys, int B, long long v_local, TpEx ex) {
  const uint32_t seq = tp_seq(ex);
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const unsigned long long k = keys[b];
  const unsigned gidx = argmax_key_index(k) + (unsigned)(v_local * ex.rank);
  const unsigned long long g = (k & 0xffffffff00000000ull) | (unsigned long long)(0xffffffffu - gidx);
  for (int p = 0; p < ex.tp; ++p) {
    char* slot = tp_slot(ex, p, seq, ex.rank);
    tp_store_word_bits(slot, 2LL * b, (uint32_t)(g & 0xffffffffull), seq);
    tp_store_word_bits(slot, 2LL * b + 1, (uint32_t)(g >> 32), seq);
  }
}

}  // namespace pg

using namespace pg;

extern "C" {

int pg_tp_begin_step(int* epoch, void* stream) {
  PG_REQUIRE(epoch != nullptr, "tp_begin_step: null epoch");
  tp_begin_step_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(epoch);
  return check_launch("tp_begin_step");
}

int pg_tp_push(const float* partial, long long n, const pg_tp_exchange* ex, void* stream) {
  PG_REQUIRE(ex && partial && n > 0, "tp_push: bad arguments");
  const TpEx tex = tp_ex_from(ex);
  PG_REQUIRE(tex.tp >= 2 && tex.tp <= TP_MAX_RANKS && n * 8 <= tex.slot_bytes, "tp_push: %lld words do not fit the slot", n);
  const int grid = (int)((n + 255) / 256 < 296 ? (n + 255) / 256 : 296);
  return launch_pdl("tp_push", tp_push_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, partial, n, tex);
}

int pg_rmsnorm_reduce(void* out, void* x_out, const void* x_in, const void* w, int rows, int D, float eps,
                      const pg_tp_exchange* ex, int dtype, void* stream) {
  PG_REQUIRE(ex && rows > 0 && D % 2 == 0, "rmsnorm_reduce: bad arguments");
  const TpEx tex = tp_ex_from(ex);
  PG_REQUIRE(tex.tp >= 2 && tex.tp <= TP_MAX_RANKS && (long long)rows * D * 8 <= tex.slot_bytes,
             "rmsnorm_reduce: %d x %d words do not fit the slot", rows, D);
  PG_DISPATCH_DTYPE(dtype, T, {
    return launch_pdl("rmsnorm_reduce", rmsnorm_reduce_kernel<T>, dim3(rows), dim3(256), (size_t)D * sizeof(float),
                      (cudaStream_t)stream, (T*)out, (T*)x_out, (const T*)x_in, (const T*)w, D, eps, tex);
  });
  return PG_OK;
}

int pg_tp_keys_push(const unsigned long long* keys, int B, long long v_local, const pg_tp_exchange* ex, void* stream) {
  PG_REQUIRE(ex && keys && B > 0, "tp_keys_push: bad arguments");
  const TpEx tex = tp_ex_from(ex);
  PG_REQUIRE(tex.tp >= 2 && tex.tp <= TP_MAX_RANKS && (long long)B * 16 <= tex.slot_bytes, "tp_keys_push: slot too small");
  tp_keys_push_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(keys, B, v_local, tex);
  return check_launch("tp_keys_push");
}

}  // extern "C"
