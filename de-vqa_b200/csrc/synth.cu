// Counter-based synthetic bank generator (benchmarks and tests only; see devqa_fill_synthetic).
// Every element is a pure function of (seed, global row, column), so any row sharding and any
// append batching reproduce the same bank ("generated on device per shard").
#include <cuda_fp16.h>

#include "scan.cuh"

namespace devqa {
namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27; x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}
__device__ __forceinline__ float normal_at(uint64_t seed, uint64_t row, uint32_t col) {
  const uint64_t h = mix64(mix64(seed ^ (row * 0x9e3779b97f4a7c15ull)) + col);
  const float u1 = (float(uint32_t(h >> 40)) + 1.0f) * (1.0f / 16777217.0f);   // (0, 1)
  const float u2 = float(uint32_t(h) >> 8) * (1.0f / 16777216.0f);             // [0, 1)
  return sqrtf(-2.0f * __logf(u1)) * __cosf(6.283185307179586f * u2);
}

template <typename T>
__device__ __forceinline__ void store(T* p, float v);
template <>
__device__ __forceinline__ void store<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void store<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// one warp per row
template <typename T>
__global__ void __launch_bounds__(256) fill_kernel(T* dst, int64_t n_rows, int dim, int64_t ld, uint64_t seed,
                                                   int64_t first_row, int kind) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = int64_t(gridDim.x) * (blockDim.x >> 5);
  for (int64_t r = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n_rows; r += warps) {
    const uint64_t grow = uint64_t(first_row + r);
    T* out = dst + r * ld;
    if (kind == 0) {
      float ss = 0.f;
      for (int c = lane; c < dim; c += 32) { const float x = normal_at(seed, grow, c); ss = fmaf(x, x, ss); }
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
      const float inv = rsqrtf(ss);
      for (int c = lane; c < dim; c += 32) store(out + c, normal_at(seed, grow, c) * inv);
    } else {
      const float scale = __expf(0.5f * normal_at(seed ^ 0x5ca1ab1eull, grow, 0xffffffffu));
      uint32_t oc[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) oc[j] = uint32_t(mix64(seed + 77 + j) % uint64_t(dim));
      for (int c = lane; c < dim; c += 32) {
        float x = normal_at(seed, grow, c) * scale;
        if (uint32_t(c) == oc[0] || uint32_t(c) == oc[1] || uint32_t(c) == oc[2] || uint32_t(c) == oc[3]) x *= 50.f;
        store(out + c, x);
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) convert_kernel(const float* __restrict__ src, int64_t src_ld, T* dst,
                                                      int64_t dst_ld, int64_t n_rows, int dim) {
  const int64_t total = n_rows * dim;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / dim;
    const int c = int(i - r * dim);
    store(dst + r * dst_ld + c, src[r * src_ld + c]);
  }
}

__global__ void __launch_bounds__(256) widen_kernel(const __nv_bfloat16* __restrict__ src, int64_t src_ld, float* dst,
                                                    int64_t dst_ld, int64_t n_rows, int dim) {
  const int64_t total = n_rows * dim;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / dim;
    const int c = int(i - r * dim);
    dst[r * dst_ld + c] = __bfloat162float(src[r * src_ld + c]);
  }
}

// Shadow rows: fp16 of the fp32 row scaled by a per-row power of two that puts the row's largest magnitude in
// [2^13, 2^14) -- no overflow (fp16 max 65504), and only elements below 2^-27 of the row's maximum reach the subnormal
// range (absolute error 2^-25 there: nothing against the 2^-11 relative rounding of everything else).  scale_out[r]
// = 2^-e is what the scan multiplies the row's accumulators with.  One warp per row, two passes (the second from L2).
__global__ void __launch_bounds__(256) shadow_convert_kernel(const float* __restrict__ src, int64_t src_ld, __half* dst,
                                                             int64_t dst_ld, float* __restrict__ scale_out, int64_t n_rows,
                                                             int dim) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = int64_t(gridDim.x) * (blockDim.x >> 5);
  for (int64_t r = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n_rows; r += warps) {
    const float* in = src + r * src_ld;
    float m = 0.f;
    for (int c = lane; c < dim; c += 32) m = fmaxf(m, fabsf(in[c]));      // (fmaxf drops NaN: a NaN element stays NaN below)
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    int e = 0;
    if (m > 0.f && m <= 3.0e38f) {
      int ex;
      frexpf(m, &ex);                          // m = f * 2^ex, f in [0.5, 1)  ->  m * 2^(14 - ex) in [2^13, 2^14)
      e = 14 - ex;
      if (e > 120) e = 120;                    // (rows of subnormal magnitude: ||row||^2 is 0 in fp32 anyway)
    }
    // two factors: 2^e itself may be out of float range for rows of subnormal magnitude
    const float up1 = exp2f(float(e / 2)), up2 = exp2f(float(e - e / 2));
    __half* out = dst + r * dst_ld;
    for (int c = lane; c < dim; c += 32) out[c] = __float2half_rn(in[c] * up1 * up2);
    if (lane == 0) scale_out[r] = exp2f(float(-e));                          // (e >= -113: representable or subnormal-exact)
  }
}

}  // namespace

int shadow_convert_launch(const float* src, int64_t src_ld, void* dst, int64_t dst_ld, float* scale_out, int64_t n_rows,
                          int64_t dim, cudaStream_t stream) {
  if (n_rows == 0) return DEVQA_OK;
  const int64_t want = (n_rows + 7) / 8;
  const unsigned grid = unsigned(want < 148 * 16 ? want : 148 * 16);
  shadow_convert_kernel<<<grid, 256, 0, stream>>>(src, src_ld, static_cast<__half*>(dst), dst_ld, scale_out, n_rows, int(dim));
  DEVQA_CUDA(cudaGetLastError());
  count_launch();
  return DEVQA_OK;
}

int widen_rows_launch(const void* src, int64_t src_ld, float* dst, int64_t dst_ld, int64_t n_rows, int64_t dim,
                      cudaStream_t stream) {
  if (n_rows == 0) return DEVQA_OK;
  const int64_t want = (n_rows * dim + 255) / 256;
  const unsigned grid = unsigned(want < 148 * 8 ? want : 148 * 8);
  widen_kernel<<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(src), src_ld, dst, dst_ld, n_rows, int(dim));
  DEVQA_CUDA(cudaGetLastError());
  count_launch();
  return DEVQA_OK;
}

int fill_synthetic_launch(void* dst, int dtype, int64_t n_rows, int64_t dim, int64_t ld, uint64_t seed,
                          int64_t first_global_row, int kind, cudaStream_t stream) {
  if (n_rows == 0) return DEVQA_OK;
  const int64_t want = (n_rows + 7) / 8;
  const unsigned grid = unsigned(want < 148 * 16 ? want : 148 * 16);
  if (dtype == DEVQA_F32)
    fill_kernel<float><<<grid, 256, 0, stream>>>(static_cast<float*>(dst), n_rows, int(dim), ld, seed,
                                                 first_global_row, kind);
  else
    fill_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<__nv_bfloat16*>(dst), n_rows, int(dim), ld,
                                                         seed, first_global_row, kind);
  DEVQA_CUDA(cudaGetLastError());
  count_launch();
  return DEVQA_OK;
}

int convert_rows_launch(const float* src, int64_t src_ld, void* dst, int dst_dtype, int64_t dst_ld, int64_t n_rows,
                        int64_t dim, cudaStream_t stream) {
  if (n_rows == 0) return DEVQA_OK;
  const int64_t total = n_rows * dim;
  const int64_t want = (total + 255) / 256;
  const unsigned grid = unsigned(want < 148 * 8 ? want : 148 * 8);
  if (dst_dtype == DEVQA_F32)
    convert_kernel<float><<<grid, 256, 0, stream>>>(src, src_ld, static_cast<float*>(dst), dst_ld, n_rows, int(dim));
  else
    convert_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(src, src_ld, static_cast<__nv_bfloat16*>(dst), dst_ld,
                                                            n_rows, int(dim));
  DEVQA_CUDA(cudaGetLastError());
  count_launch();
  return DEVQA_OK;
}

}  // namespace devqa
