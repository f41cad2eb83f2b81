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

// ---- element loaders: NV columns per lane per chunk ---------------------------------------
template <typename T, int NV>
struct Loader;
template <>
struct Loader<float, 4> {
  static __device__ __forceinline__ void load(const float* row, int col, float (&b)[4]) {
    const float4 t = ld_stream_f4(reinterpret_cast<const float4*>(row + col));
    b[0] = t.x; b[1] = t.y; b[2] = t.z; b[3] = t.w;
  }
};
template <>
struct Loader<float, 1> {
  static __device__ __forceinline__ void load(const float* row, int col, float (&b)[1]) {
    b[0] = ld_stream_f1(row + col);
  }
};
template <>
struct Loader<__nv_bfloat16, 8> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* row, int col, float (&b)[8]) {
    const uint4 t = ld_stream_u4(reinterpret_cast<const uint4*>(row + col));
    b[0] = __uint_as_float(t.x << 16); b[1] = __uint_as_float(t.x & 0xffff0000u);
    b[2] = __uint_as_float(t.y << 16); b[3] = __uint_as_float(t.y & 0xffff0000u);
    b[4] = __uint_as_float(t.z << 16); b[5] = __uint_as_float(t.z & 0xffff0000u);
    b[6] = __uint_as_float(t.w << 16); b[7] = __uint_as_float(t.w & 0xffff0000u);
  }
};
template <>
struct Loader<__nv_bfloat16, 1> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* row, int col, float (&b)[1]) {
    b[0] = __bfloat162float(row[col]);
  }
};

template <typename T, int NV, int QT, int R, bool NORM>
__global__ void __launch_bounds__(kScanThreads, 1) scan_stream_kernel(const ScanParams p) {
  constexpr int V = QT * R;
  static_assert(V <= 32 && (V & (V - 1)) == 0, "QT*R must be a power of two <= 32");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int dim = int(p.dim);
  const int nchunks = (dim + 32 * NV - 1) / (32 * NV);
  const int dpad = nchunks * 32 * NV;
  const int kcap = p.kcap;

  float* sq = reinterpret_cast<float*>(smem_raw);                                  // [QT][dpad]
  uint64_t* lists = reinterpret_cast<uint64_t*>(smem_raw + size_t(QT) * dpad * 4);  // [warps][QT][kcap]

  // ---- prologue: stage (and optionally normalise) this pass's queries; clear the lists --------
  for (int q = warp; q < QT; q += kScanWarps) {
    float* dst = sq + size_t(q) * dpad;
    if (q < p.nq_pass) {
      const float* src = p.queries + size_t(p.q0 + q) * p.query_ld;
      float ss = 0.f;
      for (int c = lane; c < dpad; c += 32) {
        const float x = c < dim ? src[c] : 0.f;
        dst[c] = x;
        ss = fmaf(x, x, ss);
      }
      if (p.flags & DEVQA_NORMALIZE_QUERY) {
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) ss += __shfl_xor_sync(kFull, ss, off);
        const float nrm = sqrtf(ss);
        const float inv = 1.0f / ((p.flags & DEVQA_NOEPS) ? nrm : fmaxf(nrm, kNormEps));
        __syncwarp();
        for (int c = lane; c < dpad; c += 32) dst[c] *= inv;
      }
    } else {
      for (int c = lane; c < dpad; c += 32) dst[c] = 0.f;
    }
  }
  for (int i = threadIdx.x; i < kScanWarps * QT * kcap; i += kScanThreads) lists[i] = kEmptyKey;
  __syncthreads();

  uint64_t* my_lists = lists + size_t(warp) * QT * kcap;
  const int vi = transpose_index<V>(lane);          // which (q, r) this lane ends up owning
  const int my_r = vi % R;
  const int my_q = vi / R;
  // when V < 32 several lanes end up with the same total: only the one with clear low bits reports it
  const bool q_live = my_q < p.nq_pass && (lane & ((32 >> Log2<V>::value) - 1)) == 0;
  uint64_t thr = kEmptyKey;                          // k-th key of my_q's list

  const int64_t n_groups = (p.n_rows + R - 1) / R;
  const int64_t warp_stride = int64_t(gridDim.x) * kScanWarps;
  for (int64_t g = int64_t(blockIdx.x) * kScanWarps + warp; g < n_groups; g += warp_stride) {
    const int64_t row0 = g * R;
    const T* rp[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      int64_t row = row0 + r;
      if (row >= p.n_rows) row = p.n_rows - 1;       // clamp: loads stay in bounds, result is masked
      rp[r] = static_cast<const T*>(p.bank) + row * p.bank_ld;
    }
    float v[V];
    float nr[R];
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] = 0.f;
#pragma unroll
    for (int r = 0; r < R; ++r) nr[r] = 0.f;

#pragma unroll 2
    for (int c = 0; c < nchunks; ++c) {
      const int col = (c * 32 + lane) * NV;
      float b[R][NV];
      if (col < dim) {
#pragma unroll
        for (int r = 0; r < R; ++r) Loader<T, NV>::load(rp[r], col, b[r]);
      } else {
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int e = 0; e < NV; ++e) b[r][e] = 0.f;
      }
#pragma unroll
      for (int q = 0; q < QT; ++q) {
        float qv[NV];
        if constexpr (NV >= 4) {
#pragma unroll
          for (int h = 0; h < NV / 4; ++h) {
            const float4 t = *reinterpret_cast<const float4*>(sq + size_t(q) * dpad + col + 4 * h);
            qv[4 * h] = t.x; qv[4 * h + 1] = t.y; qv[4 * h + 2] = t.z; qv[4 * h + 3] = t.w;
          }
        } else {
          qv[0] = sq[size_t(q) * dpad + col];
        }
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
    }

    float score = transpose_reduce<V>(v, lane);
    if constexpr (NORM) {
      // rows are the low bits of the value index, so the lane that owns (q, r) also owns ||row r||^2
      const float ss = transpose_reduce<R>(nr, lane);
      const float nrm = sqrtf(ss);
      score *= 1.0f / ((p.flags & DEVQA_NOEPS) ? nrm : fmaxf(nrm, kNormEps));
    }
    const int64_t row = row0 + my_r;
    const bool live = q_live && row < p.n_rows;
    const uint64_t key = live ? pack_key(score, uint32_t(p.idmap(row))) : kEmptyKey;
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
  }

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
}

template <typename T, int NV, int QT, int R>
int launch_variant(const ScanParams& p, int grid, cudaStream_t stream) {
  const int nchunks = (int(p.dim) + 32 * NV - 1) / (32 * NV);
  const size_t smem = size_t(QT) * nchunks * 32 * NV * 4 + size_t(kScanWarps) * QT * p.kcap * 8;
  const bool norm = (p.flags & DEVQA_NORMALIZE_BANK) != 0;
  auto kern = norm ? scan_stream_kernel<T, NV, QT, R, true> : scan_stream_kernel<T, NV, QT, R, false>;
  DEVQA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  kern<<<grid, kScanThreads, smem, stream>>>(p);
  DEVQA_CUDA(cudaGetLastError());
  count_launch();
  return DEVQA_OK;
}

template <typename T, int NV>
int launch_qt(const ScanParams& p, int qt, int grid, cudaStream_t stream) {
  switch (qt) {
    case 8: return launch_variant<T, NV, 8, 4>(p, grid, stream);
    case 4: return launch_variant<T, NV, 4, 8>(p, grid, stream);
    case 2: return launch_variant<T, NV, 2, 8>(p, grid, stream);
    default: return launch_variant<T, NV, 1, 8>(p, grid, stream);
  }
}

}  // namespace

size_t scan_stream_smem_bytes(int64_t dim, int nv, int qt, int kcap) {
  const int nchunks = (int(dim) + 32 * nv - 1) / (32 * nv);
  return size_t(qt) * nchunks * 32 * nv * 4 + size_t(kScanWarps) * qt * kcap * 8;
}

int scan_stream_rows_per_group(int qt) { return qt == 8 ? 4 : 8; }

// One pass: queries [q0, q0 + nq_pass) against the whole bank.  `qt` in {1,2,4,8} >= nq_pass.
int scan_stream_launch(const ScanParams& p, int bank_dtype, bool vectorised, int qt, int grid,
                       cudaStream_t stream) {
  if (bank_dtype == DEVQA_F32) {
    return vectorised ? launch_qt<float, 4>(p, qt, grid, stream) : launch_qt<float, 1>(p, qt, grid, stream);
  }
  return vectorised ? launch_qt<__nv_bfloat16, 8>(p, qt, grid, stream)
                    : launch_qt<__nv_bfloat16, 1>(p, qt, grid, stream);
}

}  // namespace devqa
