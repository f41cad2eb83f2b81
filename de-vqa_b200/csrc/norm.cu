// Stand-alone row normalisation: util.normalize_embeddings (F.normalize p=2 dim=1, eps 1e-12;
// dataset/vllm.py:66,104,117) and the no-eps x/||x|| of ike_main.py:231,238.  The search entry
// points fuse this into the scan and never need it; it exists so that the drop-in
// normalize_embeddings() of the Python shim is served by this library too.
#include "scan.cuh"

namespace devqa {
namespace {

__global__ void __launch_bounds__(256) normalize_rows_kernel(const float* __restrict__ src, int64_t src_ld, float* dst,
                                                             int64_t dst_ld, int64_t n_rows, int dim, unsigned flags) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = int64_t(gridDim.x) * (blockDim.x >> 5);
  for (int64_t r = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n_rows; r += warps) {
    const float* in = src + r * src_ld;
    float* out = dst + r * dst_ld;
    float ss = 0.f;
    for (int c = lane; c < dim; c += 32) { const float x = in[c]; ss = fmaf(x, x, ss); }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
    const float nrm = sqrtf(ss);
    const float den = (flags & DEVQA_NOEPS) ? nrm : fmaxf(nrm, kNormEps);
    for (int c = lane; c < dim; c += 32) out[c] = in[c] / den;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) row_sumsq_kernel(const T* __restrict__ rows, int64_t ld, int64_t n_rows, int dim,
                                                        float* __restrict__ out, uint32_t* max_bits) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = int64_t(gridDim.x) * (blockDim.x >> 5);
  for (int64_t r = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n_rows; r += warps) {
    const T* in = rows + r * ld;
    float ss = 0.f;
    for (int c = lane; c < dim; c += 32) { const float x = float(in[c]); ss = fmaf(x, x, ss); }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
    if (lane == 0) {
      out[r] = ss;
      if (max_bits) atomicMax(max_bits, __float_as_uint(ss));   // sums are >= 0 (or NaN, which then poisons the max on purpose)
    }
  }
}

}  // namespace

int row_sumsq_launch(const void* rows, int dtype, int64_t n_rows, int64_t dim, int64_t ld, float* out, cudaStream_t stream,
                     uint32_t* max_bits) {
  if (n_rows == 0) return DEVQA_OK;
  const int64_t want = (n_rows + 7) / 8;
  const unsigned grid = unsigned(want < 148 * 16 ? want : 148 * 16);
  if (dtype == DEVQA_F32)
    row_sumsq_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(rows), ld, n_rows, int(dim), out, max_bits);
  else
    row_sumsq_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(rows), ld, n_rows, int(dim), out, max_bits);
  DEVQA_CUDA(cudaGetLastError());
  count_launch();
  return DEVQA_OK;
}

int normalize_rows_launch(const float* src, int64_t src_ld, float* dst, int64_t dst_ld, int64_t n_rows, int64_t dim,
                          unsigned flags, cudaStream_t stream) {
  if (n_rows == 0) return DEVQA_OK;
  const int64_t want = (n_rows + 7) / 8;
  const unsigned grid = unsigned(want < 148 * 8 ? want : 148 * 8);
  normalize_rows_kernel<<<grid, 256, 0, stream>>>(src, src_ld, dst, dst_ld, n_rows, int(dim), flags);
  DEVQA_CUDA(cudaGetLastError());
  count_launch();
  return DEVQA_OK;
}

}  // namespace devqa
