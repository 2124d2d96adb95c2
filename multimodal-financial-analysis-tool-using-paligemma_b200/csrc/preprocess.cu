// Image pre-processing kernels: Pillow's 8-bit two-pass bicubic resize (libImaging/Resample.c, the arithmetic behind
// processing_paligemma.py:13-18) and the byte -> normalised CHW conversion (processing_paligemma.py:19-49).
// The coefficient tables (window start / tap count, 22-bit integer taps) come from the host
// (pg_b200/preprocess.py::resample_coeffs restates precompute_coeffs + normalize_coeffs_8bpc); the kernels do the integer
// multiply-accumulates, so the result equals Pillow's bit for bit.
#include "common.cuh"

namespace pg {

constexpr int RS_PRECISION_BITS = 32 - 8 - 2;

// One thread per output byte.  horizontal: in (rows, n_in, C) -> out (rows, n_out, C);
// vertical:   in (n_in, rows, C) -> out (n_out, rows, C).
__global__ void resample_u8_kernel(uint8_t* __restrict__ out, const uint8_t* __restrict__ in,
                                   const int32_t* __restrict__ bounds, const int32_t* __restrict__ kk, int ksize,
                                   int rows, int n_in, int n_out, int C, int vertical, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long t = i / C;
    int r, xx;
    if (vertical) { r = (int)(t % rows); xx = (int)(t / rows); }
    else { xx = (int)(t % n_out); r = (int)(t / n_out); }
    const int x0 = bounds[2 * xx], cnt = bounds[2 * xx + 1];
    const int32_t* k = kk + (size_t)xx * ksize;
    int ss = 1 << (RS_PRECISION_BITS - 1);
    if (vertical) {
      const uint8_t* p = in + ((size_t)x0 * rows + r) * C + c;
      for (int x = 0; x < cnt; ++x) ss += (int)p[(size_t)x * rows * C] * k[x];
    } else {
      const uint8_t* p = in + ((size_t)r * n_in + x0) * C + c;
      for (int x = 0; x < cnt; ++x) ss += (int)p[(size_t)x * C] * k[x];
    }
    int v = ss >> RS_PRECISION_BITS;   // clip8
    v = v < 0 ? 0 : (v > 255 ? 255 : v);
    out[i] = (uint8_t)v;
  }
}

template <typename T>
__global__ void u8_to_chw_kernel(T* __restrict__ out, const uint8_t* __restrict__ in, const float* __restrict__ lut,
                                 int H, int W, int C) {
  __shared__ float s_lut[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = lut[i];
  __syncthreads();
  const long long total = (long long)H * W * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    const int c = (int)(i / ((long long)W * H));
    out[i] = from_f<T>(s_lut[in[((size_t)y * W + x) * C + c]]);
  }
}

}  // namespace pg

using namespace pg;

extern "C" {

int pg_resample_u8(uint8_t* out, const uint8_t* in, const int32_t* bounds, const int32_t* kk, int ksize, int rows, int n_in,
                   int n_out, int channels, int vertical, void* stream) {
  PG_REQUIRE(out && in && bounds && kk && ksize > 0 && rows > 0 && n_in > 0 && n_out > 0 && channels > 0,
             "resample_u8: bad arguments");
  const long long total = (long long)rows * n_out * channels;
  int grid = (int)((total + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  resample_u8_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(out, in, bounds, kk, ksize, rows, n_in, n_out, channels,
                                                              vertical, total);
  return check_launch("resample_u8");
}

int pg_u8_to_chw(void* out, const uint8_t* in, const float* lut, int H, int W, int channels, int dtype, void* stream) {
  PG_REQUIRE(out && in && lut && H > 0 && W > 0 && channels > 0, "u8_to_chw: bad arguments");
  const long long total = (long long)H * W * channels;
  int grid = (int)((total + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  PG_DISPATCH_DTYPE(dtype, T, {
    u8_to_chw_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>((T*)out, in, lut, H, W, channels);
  });
  return check_launch("u8_to_chw");
}

}  // extern "C"
