// Cross-list top-k merge: reduces the per-CTA candidate lists of scan kernels, and the
// all-gathered per-GPU lists of the sharded search, to the final result.
//
// This is the device-side analogue of the reference's per-(query, candidate) Python heap merge
// plus final sorted() in util.semantic_search: the inputs are already
// sorted descending, keys are unique and totally ordered, so the output is independent of list
// order, grid size and GPU count.
//
// One CTA per query.  Each warp folds its share of lists into an accumulator with a bitonic
// top-k merge (registers + shuffles when the capacity is 32, shared memory otherwise); the warps'
// accumulators are then tree-merged through shared memory.
#include "scan.cuh"
#include "topk.cuh"

namespace devqa {
namespace {

constexpr int kMergeThreads = 512;
constexpr int kMergeWarps = kMergeThreads / 32;

__device__ __forceinline__ void emit(uint64_t key, int e, int top_k, int64_t q, uint64_t* out_keys,
                                     float* out_scores, int64_t* out_ids, uint32_t* kth_codes) {
  if (kth_codes) {                        // threshold seeding: only the k-th entry's score code is wanted
    if (e == top_k - 1) kth_codes[q] = uint32_t(key >> 32);
    return;
  }
  if (e >= top_k) return;
  const size_t o = size_t(q) * top_k + e;
  if (out_keys) {
    out_keys[o] = key;
  } else {
    const bool empty = key == kEmptyKey;
    out_scores[o] = empty ? -INFINITY : key_score(key);
    out_ids[o] = empty ? int64_t(-1) : int64_t(key_id(key));
  }
}

// Push this rank's list of query q to every peer, then wait for every peer's list of q (see scan.cuh, P2pExchange).
// Called by all threads of the CTA; returns (after a CTA-wide barrier) the base of the lists to merge.
__device__ __forceinline__ const uint64_t* p2p_exchange(const P2pExchange& x, int64_t q, int64_t n_queries, int list_stride) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  // every CTA reads the same number: the counter only moves when the LAST CTA of this grid leaves (p2p_leave)
  const uint32_t seq = *reinterpret_cast<volatile uint32_t*>(x.header + 1) + 1u;
  const size_t slot = seq & 1u;
  const size_t fo = slot * size_t(x.world) * kP2pMaxQ, dof = slot * size_t(x.world) * size_t(kP2pMaxKeys);
  const uint64_t* mine = x.local_keys + size_t(q) * list_stride;
  for (int peer = warp; peer < x.world; peer += nwarps) {
    uint64_t* dst = x.peer_data[peer] + dof + (size_t(x.rank) * n_queries + q) * list_stride;
    for (int e = lane; e < list_stride; e += 32) dst[e] = mine[e];
    __threadfence_system();                              // the keys are visible system-wide before the flag is
    __syncwarp();
    if (lane == 0)
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(x.peer_flags[peer] + fo + size_t(x.rank) * kP2pMaxQ + q), "r"(seq) : "memory");
  }
  for (int src = warp; src < x.world; src += nwarps) {
    if (lane == 0) {
      const uint32_t* flag = x.peer_flags[x.rank] + fo + size_t(src) * kP2pMaxQ + q;
      const long long t0 = clock64();
      uint32_t v;
      for (;;) {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (v == seq) break;
        if (clock64() - t0 > 20000000000ll) { atomicAdd(x.header, 1u); break; }   // ~10 s: a peer is gone
        __nanosleep(64);
      }
    }
    __syncwarp();
  }
  __syncthreads();
  return x.peer_data[x.rank] + dof;
}
// last statement of a CTA that exchanged: the last CTA to leave advances the sequence number
__device__ __forceinline__ void p2p_leave(const P2pExchange& x) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t seq = *reinterpret_cast<volatile uint32_t*>(x.header + 1) + 1u;
    __threadfence();
    if (atomicAdd(x.header + 2, 1u) == gridDim.x - 1) {
      x.header[2] = 0u;
      __threadfence();
      *reinterpret_cast<volatile uint32_t*>(x.header + 1) = seq;
    }
  }
}

// capacity 32: everything in registers
__global__ void __launch_bounds__(kMergeThreads) merge32_kernel(const uint64_t* __restrict__ lists, int n_lists,
                                                               int64_t n_queries, int list_stride, int list_len,
                                                               int top_k, uint64_t* out_keys, float* out_scores,
                                                               int64_t* out_ids, uint32_t* kth_codes,
                                                               const uint32_t* __restrict__ need, const P2pExchange xchg) {
  __shared__ uint64_t sacc[kMergeWarps][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t q = blockIdx.x;
  pdl_wait();
  if (need && need[q] == 0u) return;
  if (xchg.world > 0) lists = p2p_exchange(xchg, q, n_queries, list_stride);
  uint64_t acc = kEmptyKey;
  const int rev = 31 - lane;
  for (int l = warp; l < n_lists; l += kMergeWarps) {
    const uint64_t* src = lists + (size_t(l) * n_queries + q) * list_stride;
    const uint64_t b = rev < list_len ? __ldcg(src + rev) : kEmptyKey;   // (L2: peers may have written it)
    acc = warp_merge32(acc, b, lane);
  }
  sacc[warp][lane] = acc;
  __syncthreads();
  for (int s = 1; s < kMergeWarps; s <<= 1) {
    if ((warp % (2 * s)) == 0 && warp + s < kMergeWarps) {
      acc = warp_merge32(acc, sacc[warp + s][rev], lane);
    }
    __syncthreads();
    if ((warp % (2 * s)) == 0) sacc[warp][lane] = acc;
    __syncthreads();
  }
  if (warp == 0) emit(acc, lane, top_k, q, out_keys, out_scores, out_ids, kth_codes);
  if (xchg.world > 0) p2p_leave(xchg);
}

// general capacity (power of two >= 64): accumulators in shared memory
__global__ void __launch_bounds__(kMergeThreads) merge_generic_kernel(const uint64_t* __restrict__ lists,
                                                                     int n_lists, int64_t n_queries,
                                                                     int list_stride, int list_len, int kcap,
                                                                     int top_k, uint64_t* out_keys,
                                                                     float* out_scores, int64_t* out_ids,
                                                                     uint32_t* kth_codes,
                                                                     const uint32_t* __restrict__ need,
                                                                     const P2pExchange xchg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  pdl_wait();
  if (need && need[blockIdx.x] == 0u) return;
  if (xchg.world > 0) lists = p2p_exchange(xchg, blockIdx.x, n_queries, list_stride);
  const int nwarps = blockDim.x >> 5;
  uint64_t* acc = reinterpret_cast<uint64_t*>(smem_raw);       // [warps][kcap]
  uint64_t* inc = acc + size_t(nwarps) * kcap;                  // [warps][kcap]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t q = blockIdx.x;
  uint64_t* my_acc = acc + size_t(warp) * kcap;
  uint64_t* my_inc = inc + size_t(warp) * kcap;
  for (int e = lane; e < kcap; e += 32) my_acc[e] = kEmptyKey;
  __syncwarp();
  for (int l = warp; l < n_lists; l += nwarps) {
    const uint64_t* src = lists + (size_t(l) * n_queries + q) * list_stride;
    for (int e = lane; e < kcap; e += 32) my_inc[e] = e < list_len ? __ldcg(src + e) : kEmptyKey;
    __syncwarp();
    warp_list_merge(my_acc, my_inc, kcap, lane);
  }
  __syncthreads();
  for (int s = 1; s < nwarps; s <<= 1) {
    if ((warp % (2 * s)) == 0 && warp + s < nwarps)
      warp_list_merge(my_acc, acc + size_t(warp + s) * kcap, kcap, lane);
    __syncthreads();
  }
  if (warp == 0)
    for (int e = lane; e < kcap; e += 32) emit(my_acc[e], e, top_k, q, out_keys, out_scores, out_ids, kth_codes);
  if (xchg.world > 0) p2p_leave(xchg);
}

__global__ void __launch_bounds__(256) deep_emit_kernel(const uint64_t* __restrict__ pass_keys, int64_t n_queries, int kj,
                                                        int top_k, int col0, uint64_t* out_keys, float* out_scores,
                                                        int64_t* out_ids, uint64_t* below) {
  pdl_wait();
  const int64_t q = blockIdx.x;
  for (int e = threadIdx.x; e < kj; e += blockDim.x) {
    const uint64_t key = pass_keys[size_t(q) * kj + e];
    const size_t o = size_t(q) * top_k + col0 + e;
    if (out_keys) {
      out_keys[o] = key;
    } else {
      const bool empty = key == kEmptyKey;
      out_scores[o] = empty ? -INFINITY : key_score(key);
      out_ids[o] = empty ? int64_t(-1) : int64_t(key_id(key));
    }
    if (e == kj - 1) below[q] = key == kEmptyKey ? 1ull : key;      // (1: below every real key -- the bank is exhausted)
  }
}

}  // namespace

int deep_emit_launch(const uint64_t* pass_keys, int64_t n_queries, int kj, int top_k, int col0, uint64_t* out_keys,
                     float* out_scores, int64_t* out_ids, uint64_t* below, cudaStream_t stream) {
  if (n_queries == 0) return DEVQA_OK;
  DEVQA_CUDA(launch_pdl(deep_emit_kernel, dim3(unsigned(n_queries)), dim3(256), 0, stream, pass_keys, n_queries, kj, top_k, col0,
                        out_keys, out_scores, out_ids, below));
  DEVQA_CUDA(cudaGetLastError());
  count_launch();
  return DEVQA_OK;
}

int merge_lists_launch(const uint64_t* lists, int n_lists, int64_t n_queries, int list_stride,
                       int list_len, int top_k, uint64_t* out_keys, float* out_scores, int64_t* out_ids,
                       cudaStream_t stream, uint32_t* kth_codes, const uint32_t* need, const P2pExchange* xchg) {
  if (n_queries == 0) return DEVQA_OK;
  P2pExchange x{};                                           // world == 0: no exchange
  if (xchg) x = *xchg;
  const int kcap = topk_capacity(top_k > list_len ? top_k : list_len);
  if (kcap == 32) {
    DEVQA_CUDA(launch_pdl(merge32_kernel, dim3(unsigned(n_queries)), dim3(kMergeThreads), 0, stream, lists, n_lists, n_queries,
                          list_stride, list_len, top_k, out_keys, out_scores, out_ids, kth_codes, need, x));
  } else {
    const int threads = kcap <= 512 ? kMergeThreads : kMergeThreads / 2;   // keep smem <= 128 KB
    const size_t smem = size_t(2) * (threads / 32) * kcap * 8;
    DEVQA_CUDA(cudaFuncSetAttribute(merge_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    DEVQA_CUDA(launch_pdl(merge_generic_kernel, dim3(unsigned(n_queries)), dim3(threads), smem, stream, lists, n_lists, n_queries,
                          list_stride, list_len, kcap, top_k, out_keys, out_scores, out_ids, kth_codes, need, x));
  }
  DEVQA_CUDA(cudaGetLastError());
  count_launch();
  return DEVQA_OK;
}

}  // namespace devqa
