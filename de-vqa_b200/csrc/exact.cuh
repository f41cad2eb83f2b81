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
  constexpr int U = 8;                       // chunks whose loads are issued together (memory-level parallelism: the
                                             // row comes from HBM, one round trip per batch instead of one per chunk)
  const int nchunks = (dim + CH - 1) / CH;
  const bool norm = (flags & DEVQA_NORMALIZE_BANK) != 0;
  const bool want_nr = norm || row_norm != nullptr;
  float acc = 0.f, nr = 0.f;
  for (int c0 = 0; c0 < nchunks; c0 += U) {
    uint4 raw[U];                            // VEC: 16 bytes per chunk; scalar: .x holds the element's fp32 bits
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int col = (c0 + u) * CH + lane * NV;
      raw[u] = make_uint4(0u, 0u, 0u, 0u);
      if (c0 + u < nchunks && col < dim) {
        if constexpr (VEC) raw[u] = __ldg(reinterpret_cast<const uint4*>(row + col));
        else if constexpr (BF) raw[u].x = __float_as_uint(__bfloat162float(row[col]));
        else raw[u].x = __float_as_uint(__ldg(row + col));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {            // the FMAs in chunk order, exactly as the stream kernel chains them
      if (c0 + u >= nchunks) break;
      const int col = (c0 + u) * CH + lane * NV;
      float b[NV], qv[NV];
      if constexpr (VEC && BF) {
        const uint4 t = raw[u];
        b[0] = __uint_as_float(t.x << 16); b[1] = __uint_as_float(t.x & 0xffff0000u);
        b[2] = __uint_as_float(t.y << 16); b[3] = __uint_as_float(t.y & 0xffff0000u);
        b[4] = __uint_as_float(t.z << 16); b[5] = __uint_as_float(t.z & 0xffff0000u);
        b[6] = __uint_as_float(t.w << 16); b[7] = __uint_as_float(t.w & 0xffff0000u);
        const float4 u0 = *reinterpret_cast<const float4*>(sq + col), u1 = *reinterpret_cast<const float4*>(sq + col + 4);
        qv[0] = u0.x; qv[1] = u0.y; qv[2] = u0.z; qv[3] = u0.w; qv[4] = u1.x; qv[5] = u1.y; qv[6] = u1.z; qv[7] = u1.w;
      } else if constexpr (VEC) {
        const uint4 t = raw[u];
        b[0] = __uint_as_float(t.x); b[1] = __uint_as_float(t.y); b[2] = __uint_as_float(t.z); b[3] = __uint_as_float(t.w);
        const float4 qq = *reinterpret_cast<const float4*>(sq + col);
        qv[0] = qq.x; qv[1] = qq.y; qv[2] = qq.z; qv[3] = qq.w;
      } else {
        b[0] = __uint_as_float(raw[u].x);
        qv[0] = sq[col];
      }
#pragma unroll
      for (int e = 0; e < NV; ++e) acc = fmaf(b[e], qv[e], acc);
      if (want_nr) {
#pragma unroll
        for (int e = 0; e < NV; ++e) nr = fmaf(b[e], b[e], nr);
      }
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

// pull a row towards L2 ahead of its re-scoring (one 128-byte line per lane and step)
template <typename T>
__device__ __forceinline__ void prefetch_row_l2(const T* row, int dim, int lane) {
  const char* p = reinterpret_cast<const char*>(row);
  for (int off = lane * 128; off < dim * int(sizeof(T)); off += 32 * 128)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p + off));
}

}  // namespace devqa
