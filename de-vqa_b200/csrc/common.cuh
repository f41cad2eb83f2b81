// Shared device/host helpers for libdevqa (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "../../include/devqa.h"

namespace devqa {

constexpr int kWarp = 32;
constexpr float kNormEps = 1e-12f;   // F.normalize's clamp

// ---- error plumbing (host) -----------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define DEVQA_CUDA(call)                                                              \
  do {                                                                                \
    cudaError_t e__ = (call);                                                         \
    if (e__ != cudaSuccess) {                                                         \
      ::devqa::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),     \
                         __FILE__, __LINE__);                                         \
      return DEVQA_E_CUDA;                                                            \
    }                                                                                 \
  } while (0)

#define DEVQA_REQUIRE(cond, ...)             \
  do {                                       \
    if (!(cond)) {                           \
      ::devqa::set_error(__VA_ARGS__);       \
      return DEVQA_E_INVALID;                \
    }                                        \
  } while (0)

// ---- packed (score, id) keys ------------------------------------------------------------------
// key = (monotone(score) << 32) | (0xFFFFFFFF - id).  Unsigned compare on keys orders by score
// descending, then id ascending.  Any NaN maps to the maximum score code so that it ranks first,
// which is what torch.topk does in the reference.  0 = empty slot.
constexpr uint64_t kEmptyKey = 0ull;

__host__ __device__ __forceinline__ uint32_t score_code(float s) {
#ifdef __CUDA_ARCH__
  uint32_t b = __float_as_uint(s);
#else
  union { float f; uint32_t u; } cv; cv.f = s; uint32_t b = cv.u;
#endif
  if ((b & 0x7fffffffu) > 0x7f800000u) return 0xffffffffu;           // NaN -> top
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float code_score(uint32_t c) {
  uint32_t b;
  if (c == 0xffffffffu) b = 0x7fc00000u;                              // canonical NaN
  else b = (c & 0x80000000u) ? (c & 0x7fffffffu) : ~c;
#ifdef __CUDA_ARCH__
  return __uint_as_float(b);
#else
  union { float f; uint32_t u; } cv; cv.u = b; return cv.f;
#endif
}
__host__ __device__ __forceinline__ uint64_t pack_key(float s, uint32_t id) {
  return (uint64_t(score_code(s)) << 32) | uint64_t(0xffffffffu - id);
}
__host__ __device__ __forceinline__ uint32_t key_id(uint64_t k) { return 0xffffffffu - uint32_t(k); }
__host__ __device__ __forceinline__ float key_score(uint64_t k) { return code_score(uint32_t(k >> 32)); }

// local row -> global corpus id (see devqa_bank_set_id_map)
struct IdMap {
  int64_t base;
  int64_t chunk;   // 0 => contiguous
  int32_t rank;
  int32_t world;
  __host__ __device__ __forceinline__ int64_t operator()(int64_t row) const {
    if (chunk <= 0) return base + row;
    return base + ((row / chunk) * world + rank) * chunk + row % chunk;
  }
};

#ifdef __CUDACC__
// ---- programmatic dependent launch ---------------------------------------------------------------------------------
// A search is a chain of small kernels (query prep -> scan -> re-scoring -> conditional fallback -> merge); between two
// kernels of a stream the GPU idles for the ~3-5 us it takes to drain one grid and launch the next -- 40 of C2's 150 us.
// Kernels of the chain are launched with the programmatic-stream-serialization attribute: the next grid is set up
// while the previous one still runs, and every kernel starts with pdl_wait(), which blocks until the previous grid has
// completed and its memory is visible.  (Without the attribute pdl_wait() is a no-op.)  DEVQA_PDL=0 disables it.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// 128-bit streaming load: read-only path, do not allocate in L1 (the bank is touched once).
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ uint4 ld_stream_u4(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_stream_f1(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src) {
  uint32_t lo = __shfl_sync(0xffffffffu, uint32_t(v), src);
  uint32_t hi = __shfl_sync(0xffffffffu, uint32_t(v >> 32), src);
  return (uint64_t(hi) << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int m) {
  uint32_t lo = __shfl_xor_sync(0xffffffffu, uint32_t(v), m);
  uint32_t hi = __shfl_xor_sync(0xffffffffu, uint32_t(v >> 32), m);
  return (uint64_t(hi) << 32) | lo;
}
#endif

}  // namespace devqa
