// Warp-level top-k primitives on packed 64-bit keys (see common.cuh for the key format).
// Lists are sorted DESCENDING; capacity `kcap` is a power of two >= 32; empty slots hold 0.
#pragma once
#include "common.cuh"

namespace devqa {

constexpr unsigned kFull = 0xffffffffu;

__host__ __device__ inline int topk_capacity(int k) {
  int c = 32;
  while (c < k) c <<= 1;
  return c;
}

// Insert `x` (warp-uniform) into the sorted list if it ranks among the first `k` entries.
// All 32 lanes must call.  Cost for kcap = 32: 2 LDS + ballot + 1 STS.
__device__ __forceinline__ void warp_list_insert(uint64_t* list, int kcap, int k, uint64_t x, int lane) {
  int pos = 0;
  for (int j = 0; j < kcap; j += 32) pos += __popc(__ballot_sync(kFull, list[j + lane] > x));
  if (pos >= k) return;                         // warp-uniform
  for (int j = kcap - 32; j >= 0; j -= 32) {    // high blocks first: sources are still unshifted
    const int e = j + lane;
    const uint64_t prev = (e > 0) ? list[e - 1] : 0ull;
    __syncwarp();
    if (e > pos) list[e] = prev;
    else if (e == pos) list[e] = x;
    __syncwarp();
  }
}

// dst <- top-kcap of (dst U src), both sorted descending, by one warp, in shared memory.
// C[i] = max(A[i], B[kcap-1-i]) is the top-kcap set as a bitonic sequence; log2(kcap)
// compare-exchange sweeps sort it.
__device__ __forceinline__ void warp_list_merge(uint64_t* dst, const uint64_t* src, int kcap, int lane) {
  for (int e = lane; e < kcap; e += 32) {
    const uint64_t a = dst[e], b = src[kcap - 1 - e];
    dst[e] = a > b ? a : b;
  }
  __syncwarp();
  for (int stride = kcap >> 1; stride >= 1; stride >>= 1) {
    for (int i = lane; i < (kcap >> 1); i += 32) {
      const int lo = ((i / stride) * (stride << 1)) + (i % stride);
      const int hi = lo + stride;
      const uint64_t a = dst[lo], b = dst[hi];
      if (b > a) { dst[lo] = b; dst[hi] = a; }
    }
    __syncwarp();
  }
}

// Register form for kcap == 32: lane i holds A[i] (descending) and b_rev = B[31 - i].
__device__ __forceinline__ uint64_t warp_merge32(uint64_t a, uint64_t b_rev, int lane) {
  uint64_t c = a > b_rev ? a : b_rev;
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    const uint64_t o = shfl_xor_u64(c, m);
    const bool upper = (lane & m) != 0;
    c = (upper == (c > o)) ? o : c;             // lower lane keeps max, upper lane keeps min
  }
  return c;
}

// Bitonic sort of one key per lane, descending (lane 0 ends with the largest key).
__device__ __forceinline__ uint64_t warp_sort32_desc(uint64_t v, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const uint64_t o = shfl_xor_u64(v, j);
      const bool keep_max = ((lane & j) == 0) == ((lane & k) == 0);
      v = (keep_max == (v > o)) ? v : o;
    }
  }
  return v;
}

// Final result slot e of query q: packed key, or (score, id) with -inf / -1 for an empty slot.
__device__ __forceinline__ void emit_result(uint64_t key, int e, int top_k, int64_t q, uint64_t* out_keys, float* out_scores,
                                            int64_t* out_ids) {
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

}  // namespace devqa
