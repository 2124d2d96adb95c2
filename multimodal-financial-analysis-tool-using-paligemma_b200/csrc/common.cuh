// Shared device helpers for the pg_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <type_traits>

#include "../../include/pg_b200.h"

namespace pg {

typedef __nv_bfloat16 bf16;
typedef __half f16;

void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define PG_REQUIRE(cond, ...)            \
  do {                                   \
    if (!(cond)) {                       \
      pg::set_error(__VA_ARGS__);        \
      return (int)PG_ERR_INVALID;        \
    }                                    \
  } while (0)

// Dispatch a templated launcher on the runtime dtype enum.
#define PG_DISPATCH_DTYPE(dtype, T, ...)                                  \
  switch (dtype) {                                                        \
    case PG_F32: { typedef float T; __VA_ARGS__; } break;                 \
    case PG_BF16: { typedef pg::bf16 T; __VA_ARGS__; } break;             \
    case PG_F16: { typedef pg::f16 T; __VA_ARGS__; } break;               \
    default: pg::set_error("bad dtype %d", (int)(dtype)); return PG_ERR_INVALID; \
  }

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------- scalar conversion
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f<f16>(f16 v) { return __half2float(v); }

template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ f16 from_f<f16>(float v) { return __float2half_rn(v); }

// Round an fp32 value to the model dtype and back: the reference materialises a tensor of
// the model dtype after every ATen op, so fused kernels round at the same points.
template <typename T> __device__ __forceinline__ float rnd(float v) { return to_f<T>(from_f<T>(v)); }
template <> __device__ __forceinline__ float rnd<float>(float v) { return v; }
// Two values at once: ONE packed conversion (F2FP, ALU pipe) instead of two F2F conversions, which issue on the
// XU pipe (16/clk/SM) next to MUFU.EX2 / MUFU.TANH -- same round-to-nearest-even results.
template <typename T> __device__ __forceinline__ void rnd2(float& a, float& b) { a = rnd<T>(a); b = rnd<T>(b); }
template <> __device__ __forceinline__ void rnd2<bf16>(float& a, float& b) {
  uint32_t u;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(b), "f"(a));  // first source -> upper half
  a = __uint_as_float(u << 16);
  b = __uint_as_float(u & 0xffff0000u);
}
template <> __device__ __forceinline__ void rnd2<f16>(float& a, float& b) {
  const float2 f = __half22float2(__floats2half2_rn(a, b));
  a = f.x;
  b = f.y;
}

// ---------------------------------------------------------------- 16-byte vectors
template <typename T> struct Vec;  // VEC elements of T in one 128-bit access
template <> struct Vec<float> { static constexpr int N = 4; };
template <> struct Vec<bf16> { static constexpr int N = 8; };
template <> struct Vec<f16> { static constexpr int N = 8; };

__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  // weights are read exactly once per step: no L1 allocation, first to leave L2 (so they do not
  // push out the lines the prefetch chain staged for the next kernels)
  uint64_t pol;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
  return r;
}

// ---------------------------------------------------------------- L2 prefetch chain
// Every decode kernel starts by asking the L2 for its CTA's share of the NEXT kernel's weights
// (cp.async.bulk.prefetch.L2): the weights never depend on activations, so HBM keeps streaming
// through the small latency-bound kernels and across kernel boundaries.
struct Prefetch { const char* ptr; unsigned long long bytes; };
Prefetch take_prefetch();  // host side: the region registered by pg_set_next_prefetch (cleared on read)

__device__ __forceinline__ void l2_prefetch_slice(Prefetch pf) {
  if (pf.bytes == 0 || threadIdx.x != 0) return;
  const unsigned long long nblk = (unsigned long long)gridDim.x * gridDim.y * gridDim.z;
  const unsigned long long bid = blockIdx.x + (unsigned long long)gridDim.x * (blockIdx.y + (unsigned long long)gridDim.y * blockIdx.z);
  unsigned long long per = (pf.bytes + nblk - 1) / nblk;
  per = (per + 127ull) & ~127ull;
  const unsigned long long off = bid * per;
  if (off >= pf.bytes) return;
  unsigned long long n = pf.bytes - off;
  n = (n < per ? n : per) & ~15ull;
  if (n == 0) return;
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(pf.ptr + off), "r"((unsigned)n) : "memory");
}
// ---------------------------------------------------------------- programmatic dependent launch
// A kernel launched with launch_pdl() may begin while its stream predecessor is still running.
// Everything that does not depend on the predecessor (weight loads, L2 prefetch, address math)
// goes before pdl_wait(); pdl_wait() returns once the predecessor has completed and its writes
// are visible.  Every kernel of the decode chain calls both, so completion stays transitive.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }

int env_int(const char* name, int dflt);

template <typename... KArgs, typename... Args>
static int launch_pdl(const char* what, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                      cudaStream_t st, Args... args) {
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    set_error("%s: cannot reserve %zu B of shared memory", what, smem);
    cudaGetLastError();
    return PG_ERR_CUDA;
  }
  static const int pdl = env_int("PG_PDL", 1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
  if (e != cudaSuccess) {
    set_error("%s launch: %s", what, cudaGetErrorString(e));
    cudaGetLastError();
    return PG_ERR_CUDA;
  }
  return check_launch(what);
}

// Tensor-core kernels (prefill / vision chain): optional thread-block cluster + programmatic dependent launch
// Kernels launched through this MUST execute pdl_wait() in every thread
// that reads or writes global memory, and unconditionally in at least one thread per CTA.
template <typename... KArgs, typename... Args>
static int launch_tc(const char* what, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, int cluster,
                     bool small, cudaStream_t st, Args... args) {
  // measured (tools/prefill_profile.py): the early launch wins 4-6 % on the single-wave kernels of the batch-1
  // vision tower and the 260-token prefill, and loses 3 % on the many-wave batch-64 kernels: PG_TC_PDL=1 (default)
  // enables it for small launches only, 2 everywhere, 0 nowhere
  static const int pdl_mode = env_int("PG_TC_PDL", 1);
  const bool pdl = pdl_mode == 2 || (pdl_mode == 1 && small);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
  if (e != cudaSuccess) {
    set_error("%s launch: %s", what, cudaGetErrorString(e));
    cudaGetLastError();
    return PG_ERR_CUDA;
  }
  return check_launch(what);
}

__device__ __forceinline__ uint4 ldg_cached(const void* p) {
  return __ldg(reinterpret_cast<const uint4*>(p));
}

template <typename T> __device__ __forceinline__ void unpack(const uint4& u, float* f);
template <> __device__ __forceinline__ void unpack<float>(const uint4& u, float* f) {
  f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y);
  f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
}
template <> __device__ __forceinline__ void unpack<bf16>(const uint4& u, float* f) {
  // bf16 -> fp32 is a 16-bit shift
  f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
  f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
  f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xffff0000u);
  f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xffff0000u);
}
template <> __device__ __forceinline__ void unpack<f16>(const uint4& u, float* f) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 t = __half22float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}

template <typename T> __device__ __forceinline__ uint4 pack(const float* f);
template <> __device__ __forceinline__ uint4 pack<float>(const float* f) {
  return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
}
template <> __device__ __forceinline__ uint4 pack<bf16>(const float* f) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}
template <> __device__ __forceinline__ uint4 pack<f16>(const float* f) {
  uint4 u;
  __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  return u;
}

// ---------------------------------------------------------------- reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// Block-wide sum; `red` is >= 32 floats of shared memory. Every thread gets the result.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : 0.f;
  return warp_sum(t);
}

// ---------------------------------------------------------------- math
// tanh-approximate GELU exactly as ATen's CPU kernel evaluates it in fp32:
// 0.5*x*(1+tanh(sqrt(2/pi)*(x+0.044715*x^3)))  (reference modeling_gemma.py:134, modeling_siglip.py:162)
__device__ __forceinline__ float gelu_tanh(float x) {
  const float kBeta = 0.7978845608028654f;  // sqrt(2)*2/sqrt(pi)*0.5
  const float kKappa = 0.044715f;
  float inner = kBeta * (x + kKappa * (x * x * x));
  return 0.5f * x * (1.f + tanhf(inner));
}

// Same formula with the hardware tanh (MUFU.TANH, rel. error ~2^-11): used only by the 16-bit tensor-core
// epilogues, whose outputs are rounded to 8-11 mantissa bits anyway; the fp32 verification path keeps tanhf.
__device__ __forceinline__ float gelu_tanh_fast(float x) {
  const float inner = 0.7978845608028654f * (x + 0.044715f * (x * x * x));
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(inner));
  return 0.5f * x * (1.f + t);
}

// Orderable 64-bit key for (value, index) argmax: larger value wins, ties go to the LOWER index
// (torch.argmax returns the first maximal element).
__device__ __forceinline__ unsigned long long argmax_key(float v, unsigned idx) {
  unsigned b = __float_as_uint(v);
  b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  return ((unsigned long long)b << 32) | (unsigned long long)(0xffffffffu - idx);
}
__device__ __forceinline__ unsigned argmax_key_index(unsigned long long k) {
  return 0xffffffffu - (unsigned)(k & 0xffffffffull);
}
__device__ __forceinline__ float argmax_key_value(unsigned long long k) {
  unsigned b = (unsigned)(k >> 32);
  b = (b & 0x80000000u) ? (b & 0x7fffffffu) : ~b;
  return __uint_as_float(b);
}

}  // namespace pg
