// The exact fp32 score of one (row, query) pair by ONE WARP, in scan_stream.cu's arithmetic: lane l owns columns
// 128c + 4l .. + 3 of chunk c (32c + l for unaligned dims; 256c + 8l .. + 7, low half first, for bf16 rows), sequential
// FMAs over chunks, xor-butterfly 16, 8, 4, 2, 1 across lanes, then * 1/max(||row||, eps) with ||row||^2 accumulated in
// the same order.  Whoever re-scores a row with this function (refine.cu, select.cu) returns bit for bit what the
// stream kernel returns for it.
#pragma once
#include "scan.cuh"
#include "topk.cuh"

namespace devqa {

// T = float: VEC = 4 columns per lane and chunk (128-bit loads), else 1.  T = __nv_bfloat16: VEC = 8 columns per lane
// and chunk, unpacked low half first -- scan_stream.cu's Loader<T, NV> orders.
// row_norm (optional): receives ||row|| (accumulated in the same order) -- the canary below needs it for un-normalised scores
template <typename T, bool VEC>
__device__ __forceinline__ float exact_score(const T* __restrict__ row, const float* sq, int dim, unsigned flags, int lane,
                                             float* row_norm = nullptr) {
  constexpr bool BF = sizeof(T) == 2;
  constexpr int NV = VEC ? (BF ? 8 : 4) : 1;
  constexpr int CH = 32 * NV;
  const int nchunks = (dim + CH - 1) / CH;
  const bool norm = (flags & DEVQA_NORMALIZE_BANK) != 0;
  const bool want_nr = norm || row_norm != nullptr;
  float acc = 0.f, nr = 0.f;
  for (int c = 0; c < nchunks; ++c) {
    const int col = c * CH + lane * NV;
    float b[NV], qv[NV];
    if constexpr (VEC && BF) {
      uint4 t = make_uint4(0u, 0u, 0u, 0u);
      if (col < dim) t = __ldg(reinterpret_cast<const uint4*>(row + col));
      b[0] = __uint_as_float(t.x << 16); b[1] = __uint_as_float(t.x & 0xffff0000u);
      b[2] = __uint_as_float(t.y << 16); b[3] = __uint_as_float(t.y & 0xffff0000u);
      b[4] = __uint_as_float(t.z << 16); b[5] = __uint_as_float(t.z & 0xffff0000u);
      b[6] = __uint_as_float(t.w << 16); b[7] = __uint_as_float(t.w & 0xffff0000u);
      const float4 u0 = *reinterpret_cast<const float4*>(sq + col), u1 = *reinterpret_cast<const float4*>(sq + col + 4);
      qv[0] = u0.x; qv[1] = u0.y; qv[2] = u0.z; qv[3] = u0.w; qv[4] = u1.x; qv[5] = u1.y; qv[6] = u1.z; qv[7] = u1.w;
    } else if constexpr (VEC) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (col < dim) t = __ldg(reinterpret_cast<const float4*>(row + col));
      b[0] = t.x; b[1] = t.y; b[2] = t.z; b[3] = t.w;
      const float4 u = *reinterpret_cast<const float4*>(sq + col);
      qv[0] = u.x; qv[1] = u.y; qv[2] = u.z; qv[3] = u.w;
    } else {
      if constexpr (BF) b[0] = col < dim ? __bfloat162float(row[col]) : 0.f;
      else b[0] = col < dim ? __ldg(row + col) : 0.f;
      qv[0] = sq[col];
    }
#pragma unroll
    for (int e = 0; e < NV; ++e) acc = fmaf(b[e], qv[e], acc);
    if (want_nr) {
#pragma unroll
      for (int e = 0; e < NV; ++e) nr = fmaf(b[e], b[e], nr);
    }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) acc = acc + __shfl_xor_sync(kFull, acc, off);
  if (want_nr) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) nr = nr + __shfl_xor_sync(kFull, nr, off);
    const float nrm = sqrtf(nr);
    if (norm) acc *= 1.0f / ((flags & DEVQA_NOEPS) ? nrm : fmaxf(nrm, kNormEps));
    if (row_norm) *row_norm = nrm;
  }
  return acc;
}


}  // namespace devqa
