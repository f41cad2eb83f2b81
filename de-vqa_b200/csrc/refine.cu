// Exact re-scoring of the tensor-core filter's candidates (scan_tc.cu, kTf32x1).
//
// The filter scan computes, per (row, query), an approximate score s~ (one TF32 product: bank words truncated
// to 10 mantissa bits, queries rounded to TF32) and works on the upper bound  u = s~ + E >= s,
// E = tc_filter_relerr(dim) * ||row|| * ||query||  (Cauchy-Schwarz on the per-element rounding errors).
// What reaches this kernel, per query: the merged list of the K' candidates with the largest u, and the
// per-CTA lists it was merged from.  Three facts make the result exact:
//   * a row the scan dropped against a shared threshold had u < T <= exact k-th best score (T = some
//     list's k-th best u minus 2E: those k rows all have s >= u - 2E);
//   * a row that a FULL list pushed out or turned away has s <= u <= that list's last entry at the time; the
//     scan records the largest such bound per (CTA, query) (lostu), and the merged list's last entry bounds
//     what the merge dropped;
//   * a candidate that is not re-scored here has u < the k-th best exact score found so far.
// So after re-scoring, with s*_k = the k-th best exact score among the candidates: if every such bound
// is < s*_k, no lost row can enter the top-k and the answer is exact.  Otherwise (heavily duplicated or
// clustered rows, NaN) need[q] is raised and api.cu redoes the query with the fp32-exact 3xTF32 scan.
// Similarity scores still never round-trip through HBM; only K' keys per query do.
//
// The exact score is computed the way scan_stream.cu computes it -- lane l owns columns 128c + 4l .. + 3 of
// chunk c (32c + l for unaligned dims), sequential FMAs over chunks, xor-butterfly 16, 8, 4, 2, 1 across lanes,
// then * 1/max(||row||, eps) with ||row||^2 accumulated in the same order -- so both kernels return
// bit-identical scores and the kernel choice is invisible in results.
//
// One CTA per query: 8 warps re-score the candidates, then the CTA bitonic-sorts the exact keys.
#include "scan.cuh"
#include "topk.cuh"
#include "exact.cuh"

namespace devqa {
namespace {

constexpr int kRefineThreads = 256;
constexpr int kMaxKp = 256;        // refine_kernel: merged candidates per query (K' = 256 serves top_k <= 128)
constexpr int kListKp = 512;       // refine_lists_kernel: candidates per query (top_k <= 64; the bf16-shadow filter's
                                   // wide error band leaves ~200 of them on a 10M-row bank)

// Canary of the filter's error bound.  Every candidate that is re-scored has both its upper bound u (the score field of
// the key the scan produced: u = s~ + E) and its exact score s.  The whole completeness proof rests on
//     u - 2E <= s <= u,      E = qerr * (||row|| for un-normalised scores, 1 otherwise);
// the bound's accumulation term models undocumented tensor-core behaviour (scan.cuh), so it is CHECKED here instead of
// trusted: a violation flags the query for the exact fallback and is counted (devqa_filter_bound_violations).
// Non-finite values carry no information about the bound and are skipped.
__device__ __forceinline__ bool bound_violated(float u, float s, float e) {
  if (!(fabsf(u) <= 3.0e38f) || !(fabsf(s) <= 3.0e38f) || !(e <= 3.0e38f)) return false;
  return s > u || (u - s) > 2.0001f * e + 1e-30f;
}

// ex[0 .. n): exact keys, unique or empty (0)  ->  srt[0 .. n) descending, empties last (n <= blockDim.x).  Rank sort:
// thread i counts the keys above its own (n broadcast reads, no barrier in between) -- ~1/10 of the 28-phase bitonic
// network this replaced, which was a third of the kernel's time.  All threads call; srt must not alias ex.
__device__ __forceinline__ void rank_sort(const uint64_t* ex, uint64_t* srt, int n) {
  const int i = threadIdx.x;
  if (i < n) srt[i] = kEmptyKey;
  __syncthreads();
  if (i < n) {
    const uint64_t key = ex[i];
    if (key != kEmptyKey) {
      int rank = 0;
#pragma unroll 8
      for (int j = 0; j < n; ++j) rank += ex[j] > key ? 1 : 0;
      srt[rank] = key;
    }
  }
  __syncthreads();
}

template <typename T, bool VEC>
__global__ void __launch_bounds__(kRefineThreads) refine_kernel(const RefineParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint64_t ex[kMaxKp];
  __shared__ uint64_t ak[kMaxKp];
  __shared__ uint64_t srt[kMaxKp];
  __shared__ float s_prelim;
  __shared__ uint32_t s_lost;                                       // max score code of any full list's last entry
  __shared__ uint32_t s_viol;                                       // candidates whose exact score left [u - 2E, u]
  constexpr int CH = VEC ? (sizeof(T) == 2 ? 256 : 128) : 32;
  const int dim = int(p.dim);
  const int dpad = (dim + CH - 1) / CH * CH;
  float* sq = reinterpret_cast<float*>(smem_raw);                  // the query exactly as the stream kernel stages it
  const int q = blockIdx.x;
  pdl_wait();
  if (p.only_if && p.only_if[q] == 0u) return;                     // conditional run (after the exact fallback)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* src = p.queries + size_t(q) * p.query_ld;
  const float inv = p.qinv[q];
  for (int c = threadIdx.x; c < dpad; c += kRefineThreads) sq[c] = c < dim ? src[c] * inv : 0.f;
  for (int c = threadIdx.x; c < kMaxKp; c += kRefineThreads) {
    ak[c] = c < p.kp ? p.akeys[size_t(q) * p.kp + c] : kEmptyKey;
    ex[c] = kEmptyKey;
  }
  if (threadIdx.x == 0) { s_lost = 0u; s_viol = 0u; }
  __syncthreads();
  const bool canary = !p.only_if && p.qerr != nullptr;              // the keys carry upper bounds of the filter scan
  const bool dot_mode = !(p.flags & DEVQA_NORMALIZE_BANK);
  const float eq = canary ? p.qerr[q] : 0.f;

  auto rescore = [&](int c) {                                       // whole warp
    const uint64_t a = ak[c];
    uint64_t key = kEmptyKey;
    if (a != kEmptyKey) {                                           // warp-uniform
      const int64_t gid = int64_t(key_id(a));
      int64_t row = gid - p.idmap.base;
      if (p.idmap.chunk > 0) row = ((row / p.idmap.chunk) / p.idmap.world) * p.idmap.chunk + row % p.idmap.chunk;
      float nrm = 1.f;
      const float s = exact_score<T, VEC>(static_cast<const T*>(p.bank) + row * p.bank_ld, sq, dim, p.flags, lane,
                                          (canary && dot_mode) ? &nrm : nullptr);
      key = pack_key(s, uint32_t(gid));
      if (canary && lane == 0 && bound_violated(key_score(a), s, eq * nrm)) atomicAdd(&s_viol, 1u);
    }
    if (lane == 0) ex[c] = key;
  };
  // phase 1: the k candidates with the largest upper bounds
  const int k1 = p.top_k < p.kp ? p.top_k : p.kp;
  for (int c = warp; c < k1; c += kRefineThreads / 32) rescore(c);
  // meanwhile: the upper bound of everything the per-CTA lists lost while full
  {
    uint32_t lost = 0u;
    for (int g = threadIdx.x; g < p.n_lists; g += kRefineThreads) {
      const uint32_t code = p.lostu[size_t(g) * p.n_queries + q];
      lost = code > lost ? code : lost;
    }
    if (lost != 0u) atomicMax(&s_lost, lost);
  }
  __syncthreads();
  if (warp == 0) {                                                  // the k-th best exact score among those k (or -inf)
    uint64_t m = ~0ull;
    for (int c = lane; c < p.top_k; c += 32) {
      const uint64_t key = c < k1 ? ex[c] : kEmptyKey;
      m = key < m ? key : m;
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) { const uint64_t o = shfl_xor_u64(m, off); m = o < m ? o : m; }
    if (lane == 0) s_prelim = m == kEmptyKey ? -INFINITY : key_score(m);
  }
  __syncthreads();
  // phase 2: every further candidate that could still beat it
  const float prelim = s_prelim;
  for (int c = k1 + warp; c < p.kp; c += kRefineThreads / 32) {
    const uint64_t a = ak[c];
    // (after the exact fallback the keys are 3xTF32 scores, not upper bounds: its few surplus rows are all re-scored)
    if (a != kEmptyKey && (p.only_if || !(key_score(a) < prelim))) rescore(c);    // warp-uniform; NaN upper bounds are re-scored
  }
  __syncthreads();
  rank_sort(ex, srt, p.kp <= 128 ? 128 : kMaxKp);                   // descending, empties last
  // nothing lost can enter the top-k?  (not asked when this run only re-expresses an exact result in the stream
  // kernel's arithmetic)
  if (threadIdx.x == 0 && !p.only_if) {
    const uint64_t kth = srt[p.top_k - 1];
    bool complete = true;
    uint32_t lost = s_lost;                                         // full per-CTA lists
    const uint64_t mlast = ak[p.kp - 1];                            // the merged list, when it is full
    if (mlast != kEmptyKey && uint32_t(mlast >> 32) > lost) lost = uint32_t(mlast >> 32);
    if (lost != 0u) {
      // rows were pushed out somewhere: they all have exact score <= code_score(lost)
      complete = kth != kEmptyKey && code_score(lost) < key_score(kth);    // false for NaN on either side
    }
    if (s_viol != 0u) {                                             // the error bound did not hold: nothing is proven
      complete = false;
      if (p.violations) atomicAdd(p.violations, s_viol);
    }
    p.need[q] = complete ? 0u : 1u;
    if (!complete && p.fallbacks) atomicAdd(p.fallbacks, 1u);
  }
  for (int e = threadIdx.x; e < p.top_k; e += kRefineThreads) {
    const uint64_t key = srt[e];
    const size_t o = size_t(q) * p.top_k + e;
    if (p.out_keys) {
      p.out_keys[o] = key;
    } else {
      const bool empty = key == kEmptyKey;
      p.out_scores[o] = empty ? -INFINITY : key_score(key);
      p.out_ids[o] = empty ? int64_t(-1) : int64_t(key_id(key));
    }
  }
}

// The same, straight from the per-CTA lists (no merge kernel in between), for scans whose CTAs exchanged bests
// (n_lists >= top_k).  T = (top_k-th largest of the CTAs' best upper bounds) - 2E is a lower bound of the exact
// k-th best score, so only keys with u >= T can matter: typically top_k plus a handful out of n_lists x kcap.
// They are compacted into shared memory, ALL re-scored, sorted.  More than 128 of them: the query is flagged.
constexpr int kRefineListThreads = 512;

template <typename T, bool VEC>
__global__ void __launch_bounds__(kRefineListThreads) refine_lists_kernel(const RefineParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int kWide = 512;                                         // first-cut capacity (= threads: one key each)
  static_assert(kWide == kRefineListThreads, "one key per thread in the second cut");
  __shared__ uint64_t ex[kListKp];
  __shared__ uint64_t ak[kWide];
  __shared__ uint32_t s_count, s_lost, s_tcode, s_viol, s_over;
  constexpr int CH = VEC ? (sizeof(T) == 2 ? 256 : 128) : 32;
  const int dim = int(p.dim);
  const int dpad = (dim + CH - 1) / CH * CH;
  float* sq = reinterpret_cast<float*>(smem_raw);
  const int q = blockIdx.x;
  pdl_wait();
  const bool dot_mode = !(p.flags & DEVQA_NORMALIZE_BANK);
  const float eq = p.qerr[q];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* src = p.queries + size_t(q) * p.query_ld;
  const float inv = p.qinv[q];
  for (int c = threadIdx.x; c < dpad; c += kRefineListThreads) sq[c] = c < dim ? src[c] * inv : 0.f;
  for (int c = threadIdx.x; c < kListKp; c += kRefineListThreads) ex[c] = kEmptyKey;
  if (threadIdx.x == 0) { s_count = 0u; s_lost = 0u; s_tcode = 0u; s_viol = 0u; s_over = 0u; }
  const bool direct = p.dcount != nullptr;                             // one-shot / emit scans: candidates are one flat array
  // ---- T from the CTAs' bests (warp 0), lost-row bounds (warp 1; direct mode: the scan lost nothing) ----
  if (warp == 0 && p.gmax != nullptr && p.n_lists >= p.top_k) {
    const uint32_t* g = p.gmax + size_t(q) * p.gmax_stride;
    uint32_t v[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) { const int c = lane + 32 * i; v[i] = c < p.n_lists ? g[c] : 0u; }
    uint32_t x = 0u;
    for (int bit = 31; bit >= 0; --bit) {                              // exact top_k-th largest
      const uint32_t cand = x | (1u << bit);
      int cnt = 0;
#pragma unroll
      for (int i = 0; i < 5; ++i) cnt += v[i] >= cand ? 1 : 0;
      cnt = __reduce_add_sync(kFull, cnt);
      if (cnt >= p.top_k) x = cand;
    }
    if (lane == 0 && x != 0u) {
      float maxn = 1.f;
      if (!(p.flags & DEVQA_NORMALIZE_BANK)) maxn = sqrtf(__uint_as_float(*p.bank_max_sumsq));
      const float t = code_score(x) - 2.0001f * p.qerr[q] * maxn;
      if (t == t) s_tcode = score_code(t);                              // NaN: no usable bound, keep everything
    }
  } else if (warp == 1 && !direct) {
    uint32_t lost = 0u;
    for (int g = lane; g < p.n_lists; g += 32) {
      const uint32_t code = p.lostu[size_t(g) * p.n_queries + q];
      lost = code > lost ? code : lost;
    }
    lost = __reduce_max_sync(kFull, lost);
    if (lane == 0) s_lost = lost;
  }
  __syncthreads();
  // ---- compact the keys that can still matter ----
  const uint32_t tcode = s_tcode;
  const int total = direct ? 0 : p.n_lists * p.cand_kcap;
  if (direct) {
    // the flat array holds every row that beat the threshold of ITS moment; only those at or above the final T matter
    const uint32_t cnt = p.dcount[q];
    const uint32_t have = cnt < uint32_t(p.dcap) ? cnt : uint32_t(p.dcap);
    for (uint32_t i = threadIdx.x; i < have; i += kRefineListThreads) {
      const uint64_t key = p.dcand[size_t(q) * p.dcap + i];
      if (uint32_t(key >> 32) >= tcode) {
        const uint32_t slot = atomicAdd(&s_count, 1u);
        if (slot < uint32_t(kWide)) ak[slot] = key;
      }
    }
    if (cnt > uint32_t(p.dcap) && threadIdx.x == 0) s_over = 1u;       // overflow: candidates are missing
  }
  if (!direct && p.cand_count) {
    // emit mode: per-CTA regions filled from the front, counts beside them.  A warp takes every 16th region; its
    // counts are requested together, then one (rarely two) coalesced reads per region.
    constexpr int kW = kRefineListThreads / 32;
    const int gi = warp + lane * kW;                                    // lane i holds the count of region warp + 16 i
    const uint32_t my_cnt = (lane < (160 + kW - 1) / kW && gi < p.n_lists) ? p.cand_count[size_t(gi) * p.n_queries + q] : 0u;
    for (int i = 0; warp + i * kW < p.n_lists; ++i) {
      const int g = warp + i * kW;
      const uint32_t cnt = __shfl_sync(kFull, my_cnt, i);
      const uint64_t* base = p.cand + (size_t(g) * p.n_queries + q) * p.cand_kcap;
      for (uint32_t e = lane; e < cnt; e += 32) {
        const uint64_t key = base[e];
        if (key != kEmptyKey && uint32_t(key >> 32) >= tcode) {
          const uint32_t slot = atomicAdd(&s_count, 1u);
          if (slot < uint32_t(kWide)) ak[slot] = key;
        }
      }
    }
  } else {
    for (int i = threadIdx.x; i < total; i += kRefineListThreads) {
      const int g = i / p.cand_kcap, e = i - g * p.cand_kcap;
      const uint64_t key = p.cand[(size_t(g) * p.n_queries + q) * p.cand_kcap + e];
      if (key != kEmptyKey && uint32_t(key >> 32) >= tcode) {
        const uint32_t slot = atomicAdd(&s_count, 1u);
        if (slot < uint32_t(kWide)) ak[slot] = key;
      }
    }
  }
  __syncthreads();
  uint32_t n_raw = s_count;
  // (A second cut on the exact k-th best upper bound of the kept keys used to follow here.  The two-phase re-scoring
  // below supersedes it: phase 2 only touches candidates whose upper bound reaches the k-th best EXACT score.)
  const int n = n_raw < uint32_t(kListKp) ? int(n_raw) : kListKp;
  // ---- exact re-scoring, two phases (like refine_kernel): the top_k candidates by upper bound first; then only those
  // whose upper bound can still beat the k-th best exact score found.  Re-scoring is the kernel's whole cost (one
  // HBM/L2 round trip of dim * 4 bytes per candidate), and typically a third of the candidates fall in phase 2. ----
  __shared__ float s_prelim;
  __shared__ uint64_t byu[kListKp];                                    // candidates sorted by upper bound
  __syncthreads();
  const int n_sort = (n + 31) & ~31;                                   // rank sorts walk n_sort slots, not 512
  if (threadIdx.x >= n && threadIdx.x < kListKp) ak[threadIdx.x] = kEmptyKey;
  __syncthreads();
  rank_sort(ak, byu, n_sort);
  auto row_of = [&](int c) {
    int64_t row = int64_t(key_id(byu[c])) - p.idmap.base;
    if (p.idmap.chunk > 0) row = ((row / p.idmap.chunk) / p.idmap.world) * p.idmap.chunk + row % p.idmap.chunk;
    return static_cast<const T*>(p.bank) + row * p.bank_ld;
  };
  auto rescore = [&](int c) {                                          // whole warp
    const uint64_t a = byu[c];
    float nrm = 1.f;
    const float s = exact_score<T, VEC>(row_of(c), sq, dim, p.flags, lane, dot_mode ? &nrm : nullptr);
    if (lane == 0) {
      ex[c] = pack_key(s, key_id(a));
      if (bound_violated(key_score(a), s, eq * nrm)) atomicAdd(&s_viol, 1u);   // canary, see bound_violated()
    }
  };
  constexpr int kWarps = kRefineListThreads / 32;
  const int k1 = p.top_k < n ? p.top_k : n;
  for (int c = warp; c < k1; c += kWarps) {
    // the candidates' rows left L2 long ago (the scan has streamed gigabytes since): pull the warp's NEXT row in
    // while this one is scored -- the kernel is a chain of memory round trips, not of flops
    if (c + kWarps < k1) prefetch_row_l2(row_of(c + kWarps), dim, lane);
    rescore(c);
  }
  __syncthreads();
  if (warp == 0) {                                                     // the k-th best exact score among those (or -inf)
    uint64_t m = ~0ull;
    for (int c = lane; c < p.top_k; c += 32) {
      const uint64_t key = c < k1 ? ex[c] : kEmptyKey;
      m = key < m ? key : m;
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) { const uint64_t o = shfl_xor_u64(m, off); m = o < m ? o : m; }
    if (lane == 0) s_prelim = m == kEmptyKey ? -INFINITY : key_score(m);
  }
  __syncthreads();
  const float prelim = s_prelim;
  for (int c = k1 + warp; c < n; c += kWarps)
    if (!(key_score(byu[c]) < prelim)) rescore(c);                     // warp-uniform; NaN upper bounds are re-scored
  __syncthreads();
  uint64_t* const srt = ak;                                            // (the compacted keys are dead now)
  rank_sort(ex, srt, n_sort);
  if (threadIdx.x == 0) {
    const uint64_t kth = srt[p.top_k - 1];
    bool complete = n_raw <= uint32_t(kListKp) && s_over == 0u;        // otherwise candidates were cut off
    const uint32_t lost = s_lost;
    if (complete && lost != 0u)
      complete = kth != kEmptyKey && code_score(lost) < key_score(kth);   // false for NaN on either side
    if (s_viol != 0u) {                                                // the error bound did not hold: nothing is proven
      complete = false;
      if (p.violations) atomicAdd(p.violations, s_viol);
    }
    p.need[q] = complete ? 0u : 1u;
    if (!complete && p.fallbacks) atomicAdd(p.fallbacks, 1u);
  }
  for (int e = threadIdx.x; e < p.top_k; e += kRefineListThreads) {
    const uint64_t key = srt[e];
    const size_t o = size_t(q) * p.top_k + e;
    if (p.out_keys) {
      p.out_keys[o] = key;
    } else {
      const bool empty = key == kEmptyKey;
      p.out_scores[o] = empty ? -INFINITY : key_score(key);
      p.out_ids[o] = empty ? int64_t(-1) : int64_t(key_id(key));
    }
  }
}

}  // namespace

// scan_stream's rule for the vectorised column order (4 fp32 / 8 bf16 columns per lane and chunk)
static bool refine_vectorised(const RefineParams& p) {
  const int nv = p.bank_dtype == DEVQA_F32 ? 4 : 8;
  return p.dim % nv == 0 && p.bank_ld % nv == 0 && reinterpret_cast<uintptr_t>(p.bank) % 16 == 0;
}
static size_t refine_query_smem(const RefineParams& p, bool vec) {
  const int ch = vec ? (p.bank_dtype == DEVQA_F32 ? 128 : 256) : 32;
  return size_t((p.dim + ch - 1) / ch * ch) * 4;
}

int refine_lists_launch(const RefineParams& p, cudaStream_t stream) {
  if (p.n_queries == 0) return DEVQA_OK;
  if (p.dcount) DEVQA_REQUIRE(p.top_k <= 64 && p.dcand && p.n_lists <= 160, "refine: unsupported direct candidate geometry");
  else DEVQA_REQUIRE(p.top_k <= 64 && p.n_lists <= 160 && p.n_lists >= p.top_k, "refine: unsupported list geometry");
  const bool vec = refine_vectorised(p);
  const size_t smem = refine_query_smem(p, vec);
  void (*kern)(RefineParams) = p.bank_dtype == DEVQA_F32
      ? (vec ? refine_lists_kernel<float, true> : refine_lists_kernel<float, false>)
      : (vec ? refine_lists_kernel<__nv_bfloat16, true> : refine_lists_kernel<__nv_bfloat16, false>);
  if (smem > 40 * 1024) DEVQA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  DEVQA_CUDA(launch_pdl(kern, dim3(unsigned(p.n_queries)), dim3(kRefineListThreads), smem, stream, p));
  DEVQA_CUDA(cudaGetLastError());
  count_launch();
  return DEVQA_OK;
}

int refine_launch(const RefineParams& p, cudaStream_t stream) {
  if (p.n_queries == 0) return DEVQA_OK;
  DEVQA_REQUIRE(p.kp <= kMaxKp && p.top_k <= p.kp, "refine: candidate list of %d for top_k %d unsupported", p.kp, p.top_k);
  const bool vec = refine_vectorised(p);
  const size_t smem = refine_query_smem(p, vec);
  void (*kern)(RefineParams) = p.bank_dtype == DEVQA_F32
      ? (vec ? refine_kernel<float, true> : refine_kernel<float, false>)
      : (vec ? refine_kernel<__nv_bfloat16, true> : refine_kernel<__nv_bfloat16, false>);
  if (smem > 48 * 1024) DEVQA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  DEVQA_CUDA(launch_pdl(kern, dim3(unsigned(p.n_queries)), dim3(kRefineThreads), smem, stream, p));
  DEVQA_CUDA(cudaGetLastError());
  count_launch();
  return DEVQA_OK;
}

}  // namespace devqa
