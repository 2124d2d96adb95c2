// Tensor-parallel exchange over NVLink peer memory, fused into the decode kernels.
//
// The tensor-parallel Gemma decoder needs the sum over ranks of the (B, D) partial outputs of o_proj and
// down_proj, 2 x 18 times per token.  Instead of a collective launch between the kernels, the PRODUCER
// (the GEMV epilogue, or pg_tp_push for the tensor-core step) stores its fp32 partial straight into every
// rank's exchange buffer over NVLink, and the CONSUMER (the RMSNorm prologue of the next kernel) finds all
// partials in its LOCAL memory, sums them in rank order (bit-identical on every rank), adds the residual and
// carries on.  No fence, no flag round trip: every 8-byte word carries its own sequence number
// ({fp32 value, seq}: the "LL" protocol), so a word is valid exactly when its flag matches.
//
// Buffer of one rank:   region | parity (seq & 1) | source rank | 8-byte words
// Sequence number:      *epoch * stride + index   (epoch: device counter bumped once per decode step by
//                       pg_tp_begin_step; index: position of the exchange inside the step, baked into the launch)
// Reuse safety: a producer can only be TWO exchanges ahead of the slowest consumer (it had to consume the
// exchange in between, which needed that consumer's rank to produce, which in stream order follows its consume),
// so two parities suffice.  Flags only ever grow; nothing is reset, so a captured CUDA graph replays forever.
// Waits are bounded by wall clock (globaltimer): a lost peer raises an error flag instead of hanging the GPU.
//
// Ranks live on different GPUs.  Tests emulate N ranks on ONE GPU by launching every rank's producer before
// any rank's consumer on a single stream (pg_b200/dist.py::LockstepGroup): no kernel ever waits for a later one.
#pragma once
#include "common.cuh"

namespace pg {

struct TpEx {
  char* const* peers;
  const int* epoch;
  int* err_dev;
  int* err_host;
  long long region_off, slot_bytes;
  int rank, tp, index, stride;
};

static inline TpEx tp_ex_from(const pg_tp_exchange* e) {
  TpEx x;
  if (!e) {
    x = TpEx{nullptr, nullptr, nullptr, nullptr, 0, 0, 0, 1, 0, 1};
    return x;
  }
  x.peers = reinterpret_cast<char* const*>(e->peers);
  x.epoch = e->epoch;
  x.err_dev = e->err_dev;
  x.err_host = e->err_host;
  x.region_off = e->region_off;
  x.slot_bytes = e->slot_bytes;
  x.rank = e->rank;
  x.tp = e->size;
  x.index = e->index;
  x.stride = e->stride;
  return x;
}

constexpr int TP_MAX_RANKS = 8;
constexpr unsigned long long TP_TIMEOUT_NS = 10ull * 1000 * 1000 * 1000;

__device__ __forceinline__ uint32_t tp_seq(const TpEx& e) {
  int ep;
  asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(ep) : "l"(e.epoch) : "memory");
  return (uint32_t)ep * (uint32_t)e.stride + (uint32_t)e.index;
}
// slot of source rank `src` inside rank `dst`'s buffer
__device__ __forceinline__ char* tp_slot(const TpEx& e, int dst, uint32_t seq, int src) {
  return e.peers[dst] + e.region_off + (long long)((int)(seq & 1u) * e.tp + src) * e.slot_bytes;
}
__device__ __forceinline__ void tp_store_word(char* slot, long long i, float v, uint32_t seq) {
  const unsigned long long w = ((unsigned long long)seq << 32) | (unsigned long long)__float_as_uint(v);
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(slot + i * 8), "l"(w) : "memory");
}
__device__ __forceinline__ void tp_store_word_bits(char* slot, long long i, uint32_t bits, uint32_t seq) {
  const unsigned long long w = ((unsigned long long)seq << 32) | (unsigned long long)bits;
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(slot + i * 8), "l"(w) : "memory");
}
__device__ __forceinline__ uint4 tp_load_pair(const char* slot, long long pair) {
  uint4 r;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(slot + pair * 16) : "memory");
  return r;
}
__device__ __forceinline__ unsigned long long tp_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ bool tp_failed(const TpEx& e) {
  int v;
  asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(e.err_dev) : "memory");
  return v != 0;
}
__device__ __forceinline__ void tp_fail(const TpEx& e) {
  *reinterpret_cast<volatile int*>(e.err_dev) = 1;
  if (e.err_host) *reinterpret_cast<volatile int*>(e.err_host) = 1;
}

// Words 2*pair and 2*pair+1 of every rank's slot in the LOCAL buffer, waited for and summed in rank order.
// `dead` is the caller's cached view of the error flag: once a wait has timed out nothing waits again.
__device__ __forceinline__ float2 tp_reduce_pair(const TpEx& e, uint32_t seq, long long pair, bool& dead) {
  uint4 v[TP_MAX_RANKS];
  const char* base = tp_slot(e, e.rank, seq, 0);
#pragma unroll
  for (int r = 0; r < TP_MAX_RANKS; ++r)
    if (r < e.tp) v[r] = tp_load_pair(base + (long long)r * e.slot_bytes, pair);
  float a = 0.f, b = 0.f;
#pragma unroll
  for (int r = 0; r < TP_MAX_RANKS; ++r) {
    if (r < e.tp) {
      if ((v[r].y != seq || v[r].w != seq) && !dead) {
        const unsigned long long t0 = tp_now();
        unsigned spins = 0;
        do {
          v[r] = tp_load_pair(base + (long long)r * e.slot_bytes, pair);
          if (v[r].y == seq && v[r].w == seq) break;
          if ((++spins & 255u) == 0u && (tp_failed(e) || tp_now() - t0 > TP_TIMEOUT_NS)) {
            tp_fail(e);
            dead = true;
            break;
          }
        } while (true);
      }
      a += __uint_as_float(v[r].x);
      b += __uint_as_float(v[r].z);
    }
  }
  return make_float2(a, b);
}

// Consumer side of the key exchange: the largest key over ranks (larger value first, then the lower global
// index: torch.argmax's tie rule on the concatenated logits).  Called by step_advance_kernel.
__device__ __forceinline__ unsigned long long tp_wait_best_key(const TpEx& ex, int b) {
  const uint32_t seq = tp_seq(ex);
  bool dead = tp_failed(ex);
  unsigned long long best = 0ull;
  for (int r = 0; r < ex.tp; ++r) {
    const char* slot = tp_slot(ex, ex.rank, seq, r);
    uint4 v = tp_load_pair(slot, b);
    if ((v.y != seq || v.w != seq) && !dead) {
      const unsigned long long t0 = tp_now();
      unsigned spins = 0;
      while (true) {
        v = tp_load_pair(slot, b);
        if (v.y == seq && v.w == seq) break;
        if ((++spins & 255u) == 0u && (tp_failed(ex) || tp_now() - t0 > TP_TIMEOUT_NS)) {
          tp_fail(ex);
          dead = true;
          break;
        }
      }
    }
    const unsigned long long k = ((unsigned long long)v.z << 32) | (unsigned long long)v.x;
    best = k > best ? k : best;
  }
  return best;
}

}  // namespace pg
