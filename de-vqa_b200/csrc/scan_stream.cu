// Bank-streaming scan: the small-query-batch kernel of the retrieval hot path.
//
// Replaces, in ONE pass over the bank:
//   F.normalize(bank) / bank/||bank||   (dataset/vllm.py:104, ike_main.py:238, multimodal_evaluate.py:410)
//   torch.mm(q, bank^T)                 (util.dot_score)
//   torch.topk + heap merge             (util.semantic_search; call sites dataset/vllm.py:68, ike_main.py:44,202,240)
//
// HBM-bound by design: every bank byte is loaded exactly once with 128-bit coalesced,
// L1-bypassing loads; a warp owns R consecutive rows at a time, lanes split the columns,
// QT queries sit in shared memory and are reused across the R rows from registers.  Row norms
// are accumulated from the same registers (fused normalise).  Scores are reduced across lanes
// with a transposing butterfly (31 shuffles for 32 (row,query) pairs) and go straight into a
// per-warp sorted top-k list in shared memory, gated by the running k-th key, so the score
// matrix never exists in memory.  Each CTA tree-merges its warps' lists and emits one sorted
// candidate list per query; merge.cu reduces those across CTAs.
//
// Determinism: a row's score depends only on (row, query) -- fixed per-lane column order and a
// fixed butterfly -- so every (QT, R) variant, grid size and GPU count yields bit-identical scores.
#include "scan.cuh"
#include "topk.cuh"

namespace devqa {

namespace {

constexpr int kScanWarps = kScanThreads / 32;

template <int N>
struct Log2 { static constexpr int value = 1 + Log2<N / 2>::value; };
template <>
struct Log2<1> { static constexpr int value = 0; };

// Reduce V (power of two <= 32) per-lane partials across the warp.  After the call, the lane
// holds the full sum of value index  i = sum_s bit(lane, 4-s) << s  (s < log2 V).
template <int V>
__device__ __forceinline__ float transpose_reduce(float (&v)[V], int lane) {
  constexpr int L = Log2<V>::value;
#pragma unroll
  for (int s = 0; s < L; ++s) {
    const int off = 16 >> s;
    const int n = V >> s;
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float keep = up ? v[2 * i + 1] : v[2 * i];
      const float send = up ? v[2 * i] : v[2 * i + 1];
      v[i] = keep + __shfl_xor_sync(kFull, send, off);
    }
  }
  float x = v[0];
#pragma unroll
  for (int off = 16 >> L; off >= 1; off >>= 1) x += __shfl_xor_sync(kFull, x, off);
  return x;
}
template <int V>
__device__ __forceinline__ int transpose_index(int lane) {
  constexpr int L = Log2<V>::value;
  int i = 0;
#pragma unroll
  for (int s = 0; s < L; ++s) i |= ((lane >> (4 - s)) & 1) << s;
  return i;
}

// ---- element loaders: NV columns per lane per chunk, kept PACKED in the register double buffer ----
template <typename T, int NV>
struct Loader;
template <>
struct Loader<float, 4> {
  using Raw = float4;
  static __device__ __forceinline__ Raw load(const float* p) { return ld_stream_f4(reinterpret_cast<const float4*>(p)); }
  static __device__ __forceinline__ Raw zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
  static __device__ __forceinline__ void unpack(const Raw& t, float (&b)[4]) { b[0] = t.x; b[1] = t.y; b[2] = t.z; b[3] = t.w; }
};
template <>
struct Loader<float, 1> {
  using Raw = float;
  static __device__ __forceinline__ Raw load(const float* p) { return ld_stream_f1(p); }
  static __device__ __forceinline__ Raw zero() { return 0.f; }
  static __device__ __forceinline__ void unpack(const Raw& t, float (&b)[1]) { b[0] = t; }
};
template <>
struct Loader<__nv_bfloat16, 8> {
  using Raw = uint4;
  static __device__ __forceinline__ Raw load(const __nv_bfloat16* p) { return ld_stream_u4(reinterpret_cast<const uint4*>(p)); }
  static __device__ __forceinline__ Raw zero() { return make_uint4(0u, 0u, 0u, 0u); }
  static __device__ __forceinline__ void unpack(const Raw& t, float (&b)[8]) {
    b[0] = __uint_as_float(t.x << 16); b[1] = __uint_as_float(t.x & 0xffff0000u);
    b[2] = __uint_as_float(t.y << 16); b[3] = __uint_as_float(t.y & 0xffff0000u);
    b[4] = __uint_as_float(t.z << 16); b[5] = __uint_as_float(t.z & 0xffff0000u);
    b[6] = __uint_as_float(t.w << 16); b[7] = __uint_as_float(t.w & 0xffff0000u);
  }
};
template <>
struct Loader<__nv_bfloat16, 1> {
  using Raw = float;
  static __device__ __forceinline__ Raw load(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ Raw zero() { return 0.f; }
  static __device__ __forceinline__ void unpack(const Raw& t, float (&b)[1]) { b[0] = t; }
};

// Shared-memory query layout: chunk-major and lane-interleaved, so that inside the chunk loop a lane
// reads its QT x (NV/4) float4 slices at compile-time offsets from ONE running pointer:
//   NV >= 4 : float4 sq4[chunk][q][h][lane]   (h < NV/4)
//   NV == 1 : float  sq [chunk][q][lane]
template <int NV, int QT>
__device__ __forceinline__ int query_slot(int col) {
  constexpr int CH = 32 * NV;
  const int c = col / CH, w = col % CH;
  const int ln = w / NV, e = w % NV;
  if constexpr (NV >= 4) return ((((c * QT) * (NV / 4) + e / 4) * 32 + ln) * 4) + (e % 4);   // + q * (NV/4)*128
  else return (c * QT) * 32 + ln;                                                            // + q * 32
}

template <typename T, int NV, int QT, int R, bool NORM>
struct ChunkMath {
  static constexpr int V = QT * R;
  using Raw = typename Loader<T, NV>::Raw;
  static __device__ __forceinline__ void load_q(const float* qp, int q, float (&qv)[NV]) {
    if constexpr (NV >= 4) {
#pragma unroll
      for (int h = 0; h < NV / 4; ++h) {
        const float4 t = *reinterpret_cast<const float4*>(qp + (q * (NV / 4) + h) * 128);
        qv[4 * h] = t.x; qv[4 * h + 1] = t.y; qv[4 * h + 2] = t.z; qv[4 * h + 3] = t.w;
      }
    } else {
      qv[0] = qp[q * 32];
    }
  }
  // v[q*R + r] += <bank chunk of row r, query q chunk>;  nr[r] += |bank chunk of row r|^2
  static __device__ __forceinline__ void apply(const Raw (&raw)[R], const float* qp, float (&v)[V], float (&nr)[R]) {
    if constexpr (R * NV <= 32) {
      // all rows unpacked once, queries streamed from shared memory
      float b[R][NV];
#pragma unroll
      for (int r = 0; r < R; ++r) Loader<T, NV>::unpack(raw[r], b[r]);
#pragma unroll
      for (int q = 0; q < QT; ++q) {
        float qv[NV];
        load_q(qp, q, qv);
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int e = 0; e < NV; ++e) v[q * R + r] = fmaf(b[r][e], qv[e], v[q * R + r]);
      }
      if constexpr (NORM) {
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int e = 0; e < NV; ++e) nr[r] = fmaf(b[r][e], b[r][e], nr[r]);
      }
    } else {
      // wide packed rows (bf16 x 8 rows): hold the few queries, unpack one row at a time
      static_assert(QT * NV <= 32, "query tile too wide for the row-outer order");
      float qv[QT][NV];
#pragma unroll
      for (int q = 0; q < QT; ++q) load_q(qp, q, qv[q]);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float b[NV];
        Loader<T, NV>::unpack(raw[r], b);
#pragma unroll
        for (int q = 0; q < QT; ++q)
#pragma unroll
          for (int e = 0; e < NV; ++e) v[q * R + r] = fmaf(b[e], qv[q][e], v[q * R + r]);
        if constexpr (NORM) {
#pragma unroll
          for (int e = 0; e < NV; ++e) nr[r] = fmaf(b[e], b[e], nr[r]);
        }
      }
    }
  }
};

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// SELECT: threshold-select epilogue (LiveEdit's hard routing, liveedit.py:132-143): instead of a top-k, every row gets a
// byte  mask[q][row] = (score * sel_scale > sel_thr[q])  -- no lists, no k cap, rows in bank order by construction.
template <typename T, int NV, int QT, int R, bool NORM, bool SELECT = false>
__global__ void __launch_bounds__(kScanThreads, 1) scan_stream_kernel(const ScanParams p) {
  constexpr int V = QT * R;
  constexpr int CH = 32 * NV;                       // columns one warp covers per chunk
  constexpr int QSTEP = NV >= 4 ? QT * (NV / 4) * 128 : QT * 32;   // floats of query smem per chunk
  static_assert(V <= 32 && (V & (V - 1)) == 0, "QT*R must be a power of two <= 32");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int dim = int(p.dim);
  const int nchunks = (dim + CH - 1) / CH;
  const int dpad = nchunks * CH;
  const int kcap = p.kcap;
  pdl_wait();

  float* sq = reinterpret_cast<float*>(smem_raw);                                   // [nchunks][QT][NV/4][32] x4
  uint64_t* lists = reinterpret_cast<uint64_t*>(smem_raw + size_t(QT) * dpad * 4);  // [warps][QT][kcap]

  // ---- prologue: stage (and optionally normalise) this pass's queries; clear the lists --------
  for (int q = warp; q < QT; q += kScanWarps) {
    float* dst = sq + (NV >= 4 ? q * (NV / 4) * 128 : q * 32);
    float inv = 0.f;                                 // dead queries are staged as zeros
    const float* src = p.queries;
    if (q < p.nq_pass) {
      src = p.queries + size_t(p.q0 + q) * p.query_ld;
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
    for (int c = lane; c < dpad; c += 32)
      dst[query_slot<NV, QT>(c)] = (c < dim && q < p.nq_pass) ? src[c] * inv : 0.f;
  }
  if constexpr (!SELECT)
    for (int i = threadIdx.x; i < kScanWarps * QT * kcap; i += kScanThreads) lists[i] = kEmptyKey;
  __syncthreads();

  uint64_t* my_lists = lists + size_t(warp) * QT * kcap;
  const int vi = transpose_index<V>(lane);          // which (q, r) this lane ends up owning
  const int my_r = vi % R;
  const int my_q = vi / R;
  // when V < 32 several lanes end up with the same total: only the one with clear low bits reports it
  const bool q_live = my_q < p.nq_pass && (lane & ((32 >> Log2<V>::value) - 1)) == 0;
  uint64_t thr = kEmptyKey;                          // k-th key of my_q's list
  uint64_t below_q = 0ull;                           // deep top-k: keys at or above this one were emitted by earlier passes
  if (p.below != nullptr && q_live) below_q = p.below[p.q0 + my_q];
  float sel_thr = 0.f;                               // SELECT: this lane's query's threshold
  uint8_t* sel_row = nullptr;                        //         and its mask row
  if constexpr (SELECT) {
    if (q_live) {
      sel_thr = p.sel_thr[p.q0 + my_q];
      sel_row = p.sel_mask + size_t(p.q0 + my_q) * p.sel_mask_ld;
    }
  }

  const int64_t n_groups = (p.n_rows + R - 1) / R;
  const int64_t warp_stride = int64_t(gridDim.x) * kScanWarps;
  const int lane_col = lane * NV;
  const T* const bank = static_cast<const T*>(p.bank);
  const float* const sq_lane = sq + (NV >= 4 ? lane * 4 : lane);

  using Raw = typename Loader<T, NV>::Raw;
  // Row groups are R consecutive rows.  The last group is shifted back so that it still reads R
  // in-bounds rows (rows it shares with the previous group are masked), which keeps all row
  // pointers of a group at a fixed stride.  The host pads banks shorter than R rows.
  const int64_t max_base = (p.n_rows > R ? p.n_rows : R) - R;
  const int64_t ld = p.bank_ld;
  int64_t g = int64_t(blockIdx.x) * kScanWarps + warp;
  const T* rp[R];                                    // lane's position in chunk 0 of each row of group g
  Raw b[R];                                          // the chunk being multiplied (register double buffer)
  int64_t base = g * R < max_base ? g * R : max_base;
#pragma unroll
  for (int r = 0; r < R; ++r) rp[r] = bank + (base + r) * ld + lane_col;
  if (g < n_groups && lane_col < dim) {
#pragma unroll
    for (int r = 0; r < R; ++r) b[r] = Loader<T, NV>::load(rp[r]);
  } else {
#pragma unroll
    for (int r = 0; r < R; ++r) b[r] = Loader<T, NV>::zero();
  }

  while (g < n_groups) {
    const int64_t gn = g + warp_stride;
    const bool has_next = gn < n_groups;
    const int64_t base_n = gn * R < max_base ? gn * R : max_base;
    if (has_next) {                                  // pull the next group's rows into L2 a whole group ahead
      const int row_bytes = dim * int(sizeof(T));
      const char* nb = reinterpret_cast<const char*>(bank + base_n * ld);
#pragma unroll
      for (int r = 0; r < R; ++r)
        for (int off = lane * 128; off < row_bytes; off += 32 * 128)
          prefetch_l2(nb + r * ld * int64_t(sizeof(T)) + off);
    }
    float v[V];
    float nr[R];
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] = 0.f;
#pragma unroll
    for (int r = 0; r < R; ++r) nr[r] = 0.f;

    const float* qp = sq_lane;
#pragma unroll 2
    for (int c = 1; c < nchunks; ++c) {              // multiply chunk c-1 while chunk c is in flight
      Raw bn[R];
      if (lane_col + c * CH < dim) {
#pragma unroll
        for (int r = 0; r < R; ++r) bn[r] = Loader<T, NV>::load(rp[r] + c * CH);
      } else {
#pragma unroll
        for (int r = 0; r < R; ++r) bn[r] = Loader<T, NV>::zero();
      }
      ChunkMath<T, NV, QT, R, NORM>::apply(b, qp, v, nr);
      qp += QSTEP;
#pragma unroll
      for (int r = 0; r < R; ++r) b[r] = bn[r];
    }
    {                                                // last chunk: chunk 0 of the NEXT group is in flight
      Raw bn[R];
#pragma unroll
      for (int r = 0; r < R; ++r) rp[r] = bank + (base_n + r) * ld + lane_col;
      if (has_next && lane_col < dim) {
#pragma unroll
        for (int r = 0; r < R; ++r) bn[r] = Loader<T, NV>::load(rp[r]);
      } else {
#pragma unroll
        for (int r = 0; r < R; ++r) bn[r] = Loader<T, NV>::zero();
      }
      ChunkMath<T, NV, QT, R, NORM>::apply(b, qp, v, nr);
#pragma unroll
      for (int r = 0; r < R; ++r) b[r] = bn[r];
    }

    float score = transpose_reduce<V>(v, lane);
    if constexpr (NORM) {
      // rows are the low bits of the value index, so the lane that owns (q, r) also owns ||row r||^2
      const float ss = transpose_reduce<R>(nr, lane);
      const float nrm = sqrtf(ss);
      score *= 1.0f / ((p.flags & DEVQA_NOEPS) ? nrm : fmaxf(nrm, kNormEps));
    }
    const int64_t row = base + my_r;
    const bool live = q_live && row >= g * R && row < p.n_rows;
    if constexpr (SELECT) {
      if (live) sel_row[row] = (score * p.sel_scale > sel_thr) ? uint8_t(1) : uint8_t(0);   // NaN: not selected, like torch's >
      base = base_n;
      g = gn;
      continue;
    }
    uint64_t key = live ? pack_key(score, uint32_t(p.idmap(row))) : kEmptyKey;
    if (below_q != 0ull && key >= below_q) key = kEmptyKey;
    unsigned hit = __ballot_sync(kFull, key > thr);
    if (hit) {
      while (hit) {
        const int src = __ffs(hit) - 1;
        hit &= hit - 1;
        const uint64_t x = shfl_u64(key, src);
        const int qx = __shfl_sync(kFull, my_q, src);
        warp_list_insert(my_lists + size_t(qx) * kcap, kcap, p.top_k, x, lane);
      }
      thr = my_lists[size_t(my_q) * kcap + (p.top_k - 1)];
    }
    base = base_n;
    g = gn;
  }

  if constexpr (SELECT) return;
  // ---- epilogue: tree-merge the warps' lists, one sorted candidate list per (CTA, query) ------
  __syncthreads();
  for (int s = 1; s < kScanWarps; s <<= 1) {
    const int pairs = kScanWarps / (2 * s);
    for (int t = warp; t < pairs * QT; t += kScanWarps) {
      const int q = t % QT;
      const int pr = t / QT;
      uint64_t* dst = lists + (size_t(pr * 2 * s) * QT + q) * kcap;
      const uint64_t* src = lists + (size_t(pr * 2 * s + s) * QT + q) * kcap;
      warp_list_merge(dst, src, kcap, lane);
    }
    __syncthreads();
  }
  for (int q = warp; q < p.nq_pass; q += kScanWarps) {
    uint64_t* out = p.cand + (size_t(blockIdx.x) * p.n_queries + (p.q0 + q)) * kcap;
    const uint64_t* src = lists + size_t(q) * kcap;
    for (int e = lane; e < kcap; e += 32) out[e] = src[e];
  }
  // ---- fused cross-CTA merge (capacity 32): the last CTA to get here folds every CTA's list ---------------
  // One launch instead of two: on the reference's one-query-at-a-time banks (6-19k rows) the second launch and
  // its ramp were a third of the device time.
  if (p.ticket != nullptr) {
    __shared__ unsigned s_ticket;
    __threadfence();                                   // this CTA's lists are visible device-wide
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = atomicAdd(p.ticket, 1u);
    __syncthreads();
    if (s_ticket != gridDim.x - 1) return;
    __threadfence();
    uint64_t* sacc = lists;                            // [kScanWarps][32], the lists are dead now
    const int rev = 31 - lane;
    for (int q = 0; q < p.nq_pass; ++q) {
      // all of this warp's lists are requested before the first merge: ONE L2 round trip instead of ten dependent
      // ones (the serial chain was half of the device time of a 6345-row search)
      constexpr int kMaxLists = (160 + kScanWarps - 1) / kScanWarps;
      uint64_t v[kMaxLists];
#pragma unroll
      for (int i = 0; i < kMaxLists; ++i) {
        const int l = warp + i * kScanWarps;
        v[i] = l < int(gridDim.x) ? __ldcg(p.cand + (size_t(l) * p.n_queries + (p.q0 + q)) * 32 + rev) : kEmptyKey;
      }
      uint64_t acc = kEmptyKey;
#pragma unroll
      for (int i = 0; i < kMaxLists; ++i)
        if (warp + i * kScanWarps < int(gridDim.x)) acc = warp_merge32(acc, v[i], lane);   // warp-uniform
      __syncthreads();
      sacc[warp * 32 + lane] = acc;
      __syncthreads();
      for (int s2 = 1; s2 < kScanWarps; s2 <<= 1) {
        if ((warp % (2 * s2)) == 0 && warp + s2 < kScanWarps) acc = warp_merge32(acc, sacc[(warp + s2) * 32 + rev], lane);
        __syncthreads();
        if ((warp % (2 * s2)) == 0) sacc[warp * 32 + lane] = acc;
        __syncthreads();
      }
      if (warp == 0) emit_result(acc, lane, p.top_k, p.q0 + q, p.out_keys, p.out_scores, p.out_ids);
    }
    if (threadIdx.x == 0) *p.ticket = 0u;              // ready for the next launch (stream-ordered)
  }
}

template <typename T, int NV, int QT, int R>
int launch_variant(const ScanParams& p, int grid, cudaStream_t stream) {
  const int nchunks = (int(p.dim) + 32 * NV - 1) / (32 * NV);
  const size_t smem = size_t(QT) * nchunks * 32 * NV * 4 + size_t(kScanWarps) * QT * p.kcap * 8;
  const bool norm = (p.flags & DEVQA_NORMALIZE_BANK) != 0;
  auto kern = norm ? scan_stream_kernel<T, NV, QT, R, true> : scan_stream_kernel<T, NV, QT, R, false>;
  DEVQA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  DEVQA_CUDA(launch_pdl(kern, dim3(grid), dim3(kScanThreads), smem, stream, p));
  DEVQA_CUDA(cudaGetLastError());
  count_launch();
  return DEVQA_OK;
}

template <typename T, int NV, int QT, int R>
int launch_select_variant(const ScanParams& p, int grid, cudaStream_t stream) {
  const int nchunks = (int(p.dim) + 32 * NV - 1) / (32 * NV);
  const size_t smem = size_t(QT) * nchunks * 32 * NV * 4;
  const bool norm = (p.flags & DEVQA_NORMALIZE_BANK) != 0;
  auto kern = norm ? scan_stream_kernel<T, NV, QT, R, true, true> : scan_stream_kernel<T, NV, QT, R, false, true>;
  DEVQA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  DEVQA_CUDA(launch_pdl(kern, dim3(grid), dim3(kScanThreads), smem, stream, p));
  DEVQA_CUDA(cudaGetLastError());
  count_launch();
  return DEVQA_OK;
}

template <typename T, int NV>
int launch_qt(const ScanParams& p, int qt, int grid, cudaStream_t stream) {
  if (p.sel_mask != nullptr)
    return qt == 4 ? launch_select_variant<T, NV, 4, 4>(p, grid, stream) : launch_select_variant<T, NV, 1, 4>(p, grid, stream);
  switch (qt) {
    case 8: return launch_variant<T, NV, 8, 4>(p, grid, stream);
    case 4: return launch_variant<T, NV, 4, 4>(p, grid, stream);
    case 2: return launch_variant<T, NV, 2, 4>(p, grid, stream);
    default: return launch_variant<T, NV, 1, 4>(p, grid, stream);
  }
}

}  // namespace

size_t scan_stream_smem_bytes(int64_t dim, int nv, int qt, int kcap) {
  const int nchunks = (int(dim) + 32 * nv - 1) / (32 * nv);
  return size_t(qt) * nchunks * 32 * nv * 4 + size_t(kScanWarps) * qt * kcap * 8;
}

int scan_stream_rows_per_group(int qt) { return 4; }

// One pass: queries [q0, q0 + nq_pass) against the whole bank.  `qt` in {1,2,4,8} >= nq_pass
// (threshold-select mode, p.sel_mask != nullptr: qt in {1,4}).
int scan_stream_launch(const ScanParams& p, int bank_dtype, bool vectorised, int qt, int grid,
                       cudaStream_t stream) {
  if (bank_dtype == DEVQA_F32) {
    return vectorised ? launch_qt<float, 4>(p, qt, grid, stream) : launch_qt<float, 1>(p, qt, grid, stream);
  }
  return vectorised ? launch_qt<__nv_bfloat16, 8>(p, qt, grid, stream)
                    : launch_qt<__nv_bfloat16, 1>(p, qt, grid, stream);
}

}  // namespace devqa
