// HICE scope gate: the RanPAC classifier that decides, per probe, whether the in-context
// lookup runs at all.
//
// Reference: our_predict_domain (easyeditor/evaluate/multimodal_evaluate.py:382-392) and HICE.run_classifier2
// (easyeditor/trainer/algs/HICE.py:238-263):
//     query_fea   = q / ||q||                                  (no eps)
//     query_fea_h = relu(query_fea @ W_rand)                   W_rand [dim, M]   (random projection, M ~ 10^4)
//     logits      = query_fea_h @ Wo.T                         Wo [n_classes, M] (ridge-regression read-out, 2 classes)
//     preds       = argmax(logits)
// and, in run_classifier2, the m-loc memory check in front of it: if the probe's best cosine similarity against the
// memory exceeds 0.81 the probe is out of scope (preds = 0) whatever the classifier says.  Up to ~11 probes per edit
// request each pay these two GEMVs plus, when in scope, a bank search (multimodal_evaluate.py:503,524,540,573).
//
// One kernel for all probes of a request: thread <-> hidden unit m (coalesced reads of W_rand rows, each word used by
// every probe), probes in shared memory, per-CTA partial logits reduced in a fixed order, the last CTA to arrive
// (ticket) sums the partials in CTA order -> deterministic logits, argmax and memory override.  W_rand (15 MB at
// M = 10^4) stays L2-resident between calls.  HBM/L2-bound integer-free fp32 work; no tensor cores wanted here.
#include "scan.cuh"
#include "topk.cuh"

namespace devqa {
namespace {

constexpr int kGateThreads = 256;
constexpr int kGateWarps = kGateThreads / 32;
constexpr int kGateQ = 16;                 // probes per pass (register accumulators)

__global__ void __launch_bounds__(kGateThreads) scope_gate_kernel(const GateParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sq = reinterpret_cast<float*>(smem_raw);                    // [kGateQ][dim] normalised probes of this pass
  __shared__ float s_red[kGateWarps][kGateQ];
  __shared__ unsigned s_ticket;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int dim = p.dim;
  const int m = blockIdx.x * kGateThreads + threadIdx.x;
  const bool m_live = m < p.m;
  pdl_wait();
  for (int q0 = 0; q0 < p.n_queries; q0 += kGateQ) {
    const int nq = min(kGateQ, p.n_queries - q0);
    __syncthreads();
    for (int q = warp; q < kGateQ; q += kGateWarps) {                // stage + normalise (scan_stream.cu's arithmetic)
      float inv = 0.f;
      const float* src = p.queries + size_t(q0 + (q < nq ? q : 0)) * p.query_ld;
      if (q < nq) {
        inv = 1.f;
        if (p.flags & DEVQA_NORMALIZE_QUERY) {
          float ss = 0.f;
          for (int c = lane; c < dim; c += 32) { const float x = src[c]; ss = fmaf(x, x, ss); }
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1) ss += __shfl_xor_sync(kFull, ss, off);
          const float nrm = sqrtf(ss);
          inv = 1.0f / ((p.flags & DEVQA_NOEPS) ? nrm : fmaxf(nrm, kNormEps));
        }
      }
      for (int c = lane; c < dim; c += 32) sq[q * dim + c] = q < nq ? src[c] * inv : 0.f;
    }
    __syncthreads();
    float h[kGateQ];
#pragma unroll
    for (int q = 0; q < kGateQ; ++q) h[q] = 0.f;
    if (m_live) {
      const float* w = p.w_rand + m;
#pragma unroll 4
      for (int d = 0; d < dim; ++d) {
        const float wv = __ldg(w + size_t(d) * p.w_ld);
#pragma unroll
        for (int q = 0; q < kGateQ; ++q) h[q] = fmaf(sq[q * dim + d], wv, h[q]);
      }
    }
#pragma unroll
    for (int q = 0; q < kGateQ; ++q) h[q] = fmaxf(h[q], 0.f);        // relu
    for (int c = 0; c < p.n_classes; ++c) {
      const float wo = m_live ? __ldg(p.wo + size_t(c) * p.wo_ld + m) : 0.f;
#pragma unroll
      for (int q = 0; q < kGateQ; ++q) {
        float v = h[q] * wo;
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(kFull, v, off);
        if (lane == 0) s_red[warp][q] = v;
      }
      __syncthreads();
      if (threadIdx.x < nq) {
        float t = 0.f;
#pragma unroll
        for (int w2 = 0; w2 < kGateWarps; ++w2) t += s_red[w2][threadIdx.x];
        p.partial[(size_t(blockIdx.x) * p.n_queries + q0 + threadIdx.x) * p.n_classes + c] = t;
      }
      __syncthreads();
    }
  }
  // ---- last CTA: partial sums in CTA order, argmax, memory override ----
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_ticket = atomicAdd(p.ticket, 1u);
  __syncthreads();
  if (s_ticket != gridDim.x - 1) return;
  __threadfence();
  for (int i = threadIdx.x; i < p.n_queries * p.n_classes; i += kGateThreads) {
    float t = 0.f;
    for (unsigned g = 0; g < gridDim.x; ++g) t += __ldcg(p.partial + size_t(g) * p.n_queries * p.n_classes + i);
    p.logits[i] = t;
  }
  __syncthreads();
  for (int q = threadIdx.x; q < p.n_queries; q += kGateThreads) {
    int best = 0;
    float bv = p.logits[size_t(q) * p.n_classes];
    for (int c = 1; c < p.n_classes; ++c) {
      const float v = p.logits[size_t(q) * p.n_classes + c];
      if (v > bv || (v != v && bv == bv)) { bv = v; best = c; }       // first maximum; NaN ranks highest, like torch.max
    }
    // HICE.py:248-250: a probe close to a remembered m-loc sample is out of scope whatever the classifier says
    if (p.mem_scores && p.mem_scores[size_t(q) * p.mem_stride] > p.mem_threshold) best = 0;
    p.preds[q] = best;
  }
  if (threadIdx.x == 0) *p.ticket = 0u;                              // ready for the next launch (stream-ordered)
}

}  // namespace

size_t scope_gate_partial_floats(const GateParams& p) {
  return size_t((p.m + kGateThreads - 1) / kGateThreads) * p.n_queries * p.n_classes;
}

int scope_gate_launch(const GateParams& p, cudaStream_t stream) {
  if (p.n_queries == 0) return DEVQA_OK;
  const size_t smem = size_t(kGateQ) * p.dim * 4;
  if (smem > 200 * 1024) {
    set_error("scope gate: dim %d does not fit shared memory", p.dim);
    return DEVQA_E_UNSUPPORTED;
  }
  if (smem > 40 * 1024) DEVQA_CUDA(cudaFuncSetAttribute(scope_gate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  const int grid = (p.m + kGateThreads - 1) / kGateThreads;
  DEVQA_CUDA(launch_pdl(scope_gate_kernel, dim3(grid), dim3(kGateThreads), smem, stream, p));
  DEVQA_CUDA(cudaGetLastError());
  count_launch();
  return DEVQA_OK;
}

}  // namespace devqa
