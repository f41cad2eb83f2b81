// Threshold-select retrieval: LiveEdit's hard routing.
//
// Reference (editor/vllm_editors/liveedit/liveedit.py:132-143): vis_sim = einsum('bed,med->bme', ivr, pool).mean(2) *
// sim_scale is a dense [1, m] similarity of the request against EVERY pool entry, retrieval_map = vis_sim > vis_sim_prot a
// boolean mask over the pool, and pool[retrieval_map] the routed experts -- no top-k anywhere.  The mean over the eqe_n
// slots of per-slot dot products is one dot product over the flattened eqe_n * module_dim vector divided by eqe_n.
//
// Here: the bank-streaming kernel (scan_stream.cu, SELECT instantiation) makes its one pass over the pool and, instead
// of feeding a top-k list, writes one byte per (query, row): score * scale > threshold[q].  Thresholds are read from
// DEVICE memory (the prototype similarity is computed on the device by the caller), so nothing synchronises with the host
// and there is no cap on how many rows may pass.  This file holds the second, small kernel: an ordered stream compaction
// of that mask into ascending row ids (the order pool[retrieval_map] yields) plus the selected rows' similarities,
// re-scored with exact.cuh -- bit for bit the scores the stream kernel compared with the threshold.
#include "exact.cuh"

namespace devqa {
namespace {

constexpr int kSelThreads = 512;
constexpr int kSelWarps = kSelThreads / 32;
constexpr int kBytesPerThread = 16;

template <typename T, bool VEC>
__global__ void __launch_bounds__(kSelThreads) select_compact_kernel(const SelectParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_warp_tot[kSelWarps];
  __shared__ float s_inv;
  constexpr int CH = VEC ? (sizeof(T) == 2 ? 256 : 128) : 32;
  const int dim = int(p.dim);
  const int dpad = (dim + CH - 1) / CH * CH;
  float* sq = reinterpret_cast<float*>(smem_raw);                    // the query exactly as the stream kernel stages it
  const int q = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* src = p.queries + size_t(q) * p.query_ld;
  pdl_wait();
  if (warp == 0) {                                                   // scan_stream.cu's prologue arithmetic
    float inv = 1.f;
    if (p.flags & DEVQA_NORMALIZE_QUERY) {
      float ss = 0.f;
      for (int c = lane; c < dim; c += 32) { const float x = src[c]; ss = fmaf(x, x, ss); }
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) ss += __shfl_xor_sync(kFull, ss, off);
      const float nrm = sqrtf(ss);
      inv = 1.0f / ((p.flags & DEVQA_NOEPS) ? nrm : fmaxf(nrm, kNormEps));
    }
    if (lane == 0) s_inv = inv;
  }
  __syncthreads();
  const float inv = s_inv;
  for (int c = threadIdx.x; c < dpad; c += kSelThreads) sq[c] = c < dim ? src[c] * inv : 0.f;

  // ---- ordered compaction: rows whose mask byte is set, ascending, the first `cap` of them ----
  const uint8_t* mask = p.mask + size_t(q) * p.mask_ld;
  int64_t* ids = p.out_ids + size_t(q) * p.cap;
  float* scores = p.out_scores + size_t(q) * p.cap;
  int64_t total = 0;                                                 // block-uniform running count
  for (int64_t base = 0; base < p.n_rows; base += int64_t(kSelThreads) * kBytesPerThread) {
    const int64_t r0 = base + int64_t(threadIdx.x) * kBytesPerThread;
    uint32_t bits = 0;
#pragma unroll
    for (int i = 0; i < kBytesPerThread; ++i)
      if (r0 + i < p.n_rows && mask[r0 + i]) bits |= 1u << i;
    const int mine = __popc(bits);
    int incl = mine;                                                 // inclusive scan inside the warp
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int o = __shfl_up_sync(kFull, incl, off);
      if (lane >= off) incl += o;
    }
    if (lane == 31) s_warp_tot[warp] = incl;
    __syncthreads();
    int before = 0, chunk_total = 0;
#pragma unroll
    for (int w = 0; w < kSelWarps; ++w) {
      const int t = s_warp_tot[w];
      if (w < warp) before += t;
      chunk_total += t;
    }
    int64_t slot = total + before + incl - mine;
    while (bits) {
      const int i = __ffs(bits) - 1;
      bits &= bits - 1;
      if (slot < p.cap) ids[slot] = r0 + i;                          // LOCAL row for now
      ++slot;
    }
    total += chunk_total;
    __syncthreads();                                                 // s_warp_tot is reused
  }
  if (threadIdx.x == 0) p.out_counts[q] = total;
  const int64_t n_sel = total < p.cap ? total : p.cap;
  for (int64_t c = n_sel + threadIdx.x; c < p.cap; c += kSelThreads) { ids[c] = -1; scores[c] = -INFINITY; }
  __syncthreads();                                                   // ids written by other threads of this CTA
  // ---- the selected rows' similarities, in the stream kernel's arithmetic ----
  for (int64_t c = warp; c < n_sel; c += kSelWarps) {
    const int64_t row = ids[c];
    const float s = exact_score<T, VEC>(static_cast<const T*>(p.bank) + row * p.bank_ld, sq, dim, p.flags, lane);
    if (lane == 0) { scores[c] = s * p.scale; ids[c] = p.idmap(row); }
  }
}

}  // namespace

int select_compact_launch(const SelectParams& p, cudaStream_t stream) {
  if (p.n_queries == 0) return DEVQA_OK;
  const int nv = p.bank_dtype == DEVQA_F32 ? 4 : 8;                   // scan_stream's rule for the vectorised column order
  const bool vec = p.dim % nv == 0 && p.bank_ld % nv == 0 && reinterpret_cast<uintptr_t>(p.bank) % 16 == 0;
  const int ch = vec ? (p.bank_dtype == DEVQA_F32 ? 128 : 256) : 32;
  const size_t smem = size_t((p.dim + ch - 1) / ch * ch) * 4;
  void (*kern)(SelectParams) = p.bank_dtype == DEVQA_F32
      ? (vec ? select_compact_kernel<float, true> : select_compact_kernel<float, false>)
      : (vec ? select_compact_kernel<__nv_bfloat16, true> : select_compact_kernel<__nv_bfloat16, false>);
  if (smem > 40 * 1024) DEVQA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  DEVQA_CUDA(launch_pdl(kern, dim3(unsigned(p.n_queries)), dim3(kSelThreads), smem, stream, p));
  DEVQA_CUDA(cudaGetLastError());
  count_launch();
  return DEVQA_OK;
}

}  // namespace devqa
