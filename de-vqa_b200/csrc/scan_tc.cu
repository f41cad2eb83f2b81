// Tensor-core scan: the large-query-batch kernel of the retrieval hot path (tcgen05 + TMEM + TMA).
//
// Same contract as scan_stream.cu (fused normalise + query x bank contraction + top-k, scores never
// reach HBM), for batches where the CUDA-core kernel turns FFMA/issue bound.
//
//   scores[row, q] = sum_d bank[row, d] * query[q, d]            M = 128 bank rows / tile, N = queries
//
// fp32 parity on tensor cores: tcgen05.mma has no fp32-input kind, kind::tf32 reads fp32 words from
// shared memory and ignores the low 13 mantissa bits.  So every operand is split x = hi + lo:
//   * queries: a prep kernel normalises them and writes hi = rna_tf32(q), lo = rna_tf32(q - hi);
//   * bank:    the raw fp32 tile IS the hi operand (the tensor core truncates it), and four "split"
//              warps (thread <-> row) compute lo = a - trunc_tf32(a) and write it straight into TENSOR
//              MEMORY (tcgen05.st; the second mma group reads its A operand from TMEM), accumulating
//              ||row||^2 from the same registers (fused normalise).  Keeping the lo tile out of shared
//              memory matters: the kernel is shared-memory-bandwidth bound (TMA writes + split reads +
//              UMMA operand reads), and this removes 32 KB of that traffic per k-block.
//   D[:, 0:Q)  = A . Qhi^T + Alo . Qhi^T        D[:, Q:2Q) = A . Qlo^T        (lo.lo is < 2^-21, dropped)
// realised as TWO mma groups per 32-column k-block: A x [Qhi;Qlo] (N = 2Q, A read once) into TMEM
// columns [0,2Q) and Alo x Qhi (N = Q) into columns [2Q,3Q).  The epilogue adds the three parts and
// multiplies by 1/||row||.
//
// Tensor cores add into the fp32 accumulator with truncation, so a long K chain into one accumulator
// drifts towards zero (measured 2e-5 relative at D = 4096 with outlier columns).  Two measures keep the
// result inside the 1e-5 contract: the small lo terms never share columns with the large hi.hi term,
// and the tensor core accumulates over at most one K-SET of 1024 columns (128 adds); K-sets of the
// same tile alternate between the two TMEM buffers and are summed in fp32 round-to-nearest in the
// epilogue threads' registers (MULTI instantiation, D > 1024).
//
// bf16 banks (BF = true): the bank is exact in bf16, so only the queries are split, three ways
// (hi + mid + lo, each bf16 = 24 mantissa bits): ONE mma group  A x [Qhi;Qmid;Qlo]  (kind::f16, N = 3Q,
// 64 columns per k-block) fills the same three column groups; the split warps only accumulate
// ||row||^2 and the second issuer idles.
//
// Warp roles (384 threads, 1 CTA/SM, persistent over 128-row tiles):
//   warp 0      TMA producer   bank tile [128 x 32] fp32 (tensor map, SWIZZLE_128B) + the query k-block
//                              [2Q x 32] fp32 as ONE contiguous bulk copy (prep kernel pre-swizzles it):
//                              the TMA engine retires ~one 128-byte box row per 5 cycles, which is the
//                              per-SM HBM share, so query rows must not go through the tensor path
//   warp 1      MMA issuer     one thread, tcgen05.mma.cta_group::1.kind::tf32, commits to mbarriers
//   warp 2      TMEM allocator (512 columns = two accumulator buffers)
//   warps 4-7   split          lo tile -> TMEM (tcgen05.st) + row norms
//   warps 8-11  epilogue         tcgen05.ld -> scale -> threshold filter -> per-query sorted lists (smem)
// Thresholds are shared across CTAs: any CTA's local k-th best score is a lower bound of the global k-th
// best, so each warp publishes its queries' k-th scores with atomicMax into a per-query global array and
// re-reads it once per tile.  Rows strictly below that bound cannot be in the result, so the filter stays
// exact while the number of list insertions drops from ~148 x k ln(n/k) to ~k ln(N/k) per query.
// Pipelines: smem ring (full_tma / full_split / empty), TMEM double buffer (tmem_full / tmem_empty).
// Small banks with many queries (C2: 64 tiles, 8 passes) would leave most SMs idle with one pass per
// launch, so a launch runs several query passes side by side: CTA c works on pass c % passes and on
// tiles c / passes, c / passes + ctas_per_pass, ...
#include <cuda.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "scan.cuh"
#include "topk.cuh"

namespace devqa {
namespace {

constexpr int kTcThreads = 384;
// Arithmetic of the contraction (template parameter MODE):
//   kTf32x3  fp32 bank, 3xTF32 split (fp32-exact scores straight from the tensor core)
//   kBf16x3  bf16 bank, queries split three ways (exact w.r.t. the bf16-rounded bank)
//   kTf32x1  fp32 bank, ONE TF32 product: an approximate score s~ with a rigorous error bound E (scan.cuh,
//            tc_filter_relerr), used as a FILTER.  The epilogue works on the UPPER BOUND u = s~ + E >= s: lists
//            are sorted by u; a CTA's k-th best u minus 2E is a lower bound T of the exact k-th best score, shared
//            across CTAs exactly like the thresholds of the other modes; rows with u < T cannot be in the result
//            and are dropped.  refine.cu re-scores the surviving candidates exactly in fp32 and proves that
//            nothing was lost (every row a full list pushed out has u <= that list's last entry).  One mma group
//            (N = queries), no split warps, no lo tile: 48 KB instead of 88 KB of shared-memory traffic per
//            k-block at 64 queries, and 128 queries per pass fit.
enum : int { kTf32x3 = 0, kBf16x3 = 1, kTf32x1 = 2, kBf16x1 = 3 };   // kBf16x1: the filter on a bf16 bank (queries rounded to bf16)
constexpr int kTileRows = 128;
constexpr int kBlockK = 32;                          // fp32 columns per k-block = one 128-byte swizzle row
constexpr int kATileBytes = kTileRows * kBlockK * 4; // 16 KB
constexpr int kNormRing = 8;
constexpr int kMaxStages = 8;
constexpr int kMaxAccBufs = 8;                      // accumulator buffers in tensor memory (2, or 512 / queries-per-pass in emit mode)
constexpr int kPrefetchAhead = 12;                  // bank k-blocks pulled into L2 ahead of the smem ring
constexpr int kSetKBlocks = 32;
// filter epilogue: hit queue entries; the same words hold the dense [16][128] score tile (8 KB), so >= 1024
__host__ __device__ constexpr int hit_cap(int qp) { return (qp > 128 || qp < 32) ? 1024 : 2048; }
// Undocumented debugging bits of `flags`.
// Bisection only -- results are WRONG when set:
constexpr unsigned kDbgNoInsert = 1u << 16, kDbgNoSplit = 1u << 17, kDbgNoMma2 = 1u << 18, kDbgNoQueryTma = 1u << 19,
                   kDbgNoMma1 = 1u << 30;
// A/B switches -- same results, another way to get them (tools/ab_epi.py, tests/test_gpu_filter.py
// test_scan_variants_are_bit_identical):
constexpr unsigned kDbgOneEpiGroup = 1u << 4,    // emit mode: one group of epilogue warps also at 128+ queries per pass
                   kDbgNoEvictFirst = 1u << 5,   // bank tiles without the L2 evict-first hint
                   kDbgNoResident = 1u << 6,     // query operand streamed with the ring stages, never resident
                   kDbgContigTiles = 1u << 7,    // TIMING ONLY (wrong results): bank tiles read as contiguous 16 KB boxes
                   kDbgL2Prefetch = 1u << 20,    // L2 prefetch of tiles ahead of the ring (round 1; slower)
                   kDbgNoBoot = 1u << 23,        // no first-tile bootstrap / 2 accumulator buffers
                   kDbgNoPair = 1u << 28,        // no CTA pairs
                   kDbgPairAll = 1u << 29;       // CTA pairs also at 64 / 128 queries per pass

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Experiment knob (DEVQA_WAIT_HINT_NS): a suspend-time hint for every mbarrier wait of the scan kernel.  0 = the
// hardware's default time limit.
__constant__ uint32_t c_wait_hint_ns = 0;
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (c_wait_hint_ns) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "HWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra HDONE_%=;\n\t"
        "bra HWAIT_%=;\n\t"
        "HDONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity), "r"(c_wait_hint_ns) : "memory");
    return;
  }
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// Bank tiles are read once per pass: evict-first keeps them from pushing the query operand (re-read by every CTA for every
// tile) and the kernel's own instructions out of L2.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_load_2d_hint(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair_hint(void* dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1), "l"(pol)
      : "memory");
}
// CTA pairs: the load lands in THIS CTA's shared memory, its bytes are counted on the LEADER's barrier (a shared::cluster
// address) -- the MMA issuer waits on one barrier for both halves of a stage, nobody has to forward anything
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
// contiguous global -> shared bulk copy (one request, no per-row tensor traversal)
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, TF32 inputs, fp32 accumulate; one thread issues for the CTA
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// bf16 inputs (K = 16 per instruction), fp32 accumulate
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// same, with the A operand read from tensor memory (lane = row, column = k)
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t addr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t addr, uint32_t (&r)[N]);
template <>
__device__ __forceinline__ void tmem_ld<4>(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
template <>
__device__ __forceinline__ void tmem_ld<8>(uint32_t addr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(addr));
}
template <>
__device__ __forceinline__ void tmem_ld<16>(uint32_t addr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(addr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void epilogue_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
// (emit mode with two column groups of epilogue warps: 256 threads)
__device__ __forceinline__ void epilogue_barrier_n(int threads) { asm volatile("bar.sync 1, %0;" ::"r"(threads) : "memory"); }

// ---- CTA-pair (cta_group::2) helpers ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait on a LOCAL barrier that a peer CTA arrives on (cluster-scope acquire)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAITC_%=:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONEC_%=;\n\t"
      "bra WAITC_%=;\n\t"
      "DONEC_%=:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// D[tmem of both CTAs] (+)= A[smem of each CTA: its 128 rows] . B[smem: N/2 rows in each CTA]^T; the leader issues
__device__ __forceinline__ void mma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// completion of all prior MMAs of the pair -> one arrival on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void mma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(uint16_t(3)) : "memory");
}
// instruction descriptor of the pair MMA: M = 256 (128 rows per CTA)
__host__ __device__ constexpr uint32_t idesc_tf32_pair(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(n >> 3) << 17) | (uint32_t(256 >> 4) << 24);
}

__host__ __device__ constexpr uint32_t idesc_bf16_pair(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(n >> 3) << 17) | (uint32_t(256 >> 4) << 24);
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (rows of 128 bytes, 8-row atoms of 1024 bytes):
// start address >> 4 | SBO = 1024 >> 4 at bit 32 | descriptor version 1 at bit 46 | layout 2 at bit 61.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t smem_addr) {
  return uint64_t((smem_addr & 0x3ffffu) >> 4) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61);
}
// The same descriptor from its two 32-bit halves: the high word is a constant, the low word is the
// 16-byte-granular start address (LBO = 0), so advancing along K is one 32-bit add.
constexpr uint32_t kDescHi = uint32_t(1024 >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint64_t desc_from_lo(uint32_t lo) { return (uint64_t(kDescHi) << 32) | lo; }
// instruction descriptor: fp32 accumulate, TF32 A and B, both K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(n >> 3) << 17) | (uint32_t(kTileRows >> 4) << 24);
}

// Position in a ring of S stages without integer division.
struct RingPos {
  int s = 0;
  uint32_t phase = 0;
  __device__ __forceinline__ void advance(int S) { if (++s == S) { s = 0; phase ^= 1u; } }
};

// the same with FP16 A and B (format field 0): the scaled-fp16 shadow of an fp32 bank
__host__ __device__ constexpr uint32_t idesc_f16(int n) {
  return (1u << 4) | (uint32_t(n >> 3) << 17) | (uint32_t(kTileRows >> 4) << 24);
}
__host__ __device__ constexpr uint32_t idesc_f16_pair(int n) {
  return (1u << 4) | (uint32_t(n >> 3) << 17) | (uint32_t(256 >> 4) << 24);
}
// instruction descriptor: fp32 accumulate, BF16 A and B, both K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t idesc_bf16(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(n >> 3) << 17) | (uint32_t(kTileRows >> 4) << 24);
}

struct TcParams {
  int64_t n_rows;
  int32_t dim;
  int32_t n_kblocks;       // ceil(dim / 32)
  int32_t n_tiles;         // ceil(n_rows / 128)
  int32_t qp;              // queries in this pass, padded to a multiple of 16 (<= 128)
  int32_t nq_pass;         // live queries in this pass
  int32_t q0;              // first query of the pass (candidate layout)
  int32_t n_queries;
  int32_t top_k, kcap;
  int32_t stages;
  int32_t passes;          // query passes running side by side in this launch (CTA c works on pass c % passes)
  uint32_t flags;
  const float* qcat;       // this pass's query operand: [n_kblocks][2*qp rows][32] fp32, rows 128B-swizzled
  uint64_t* cand;          // [grid, n_queries, kcap]
  const float* row_sumsq;  // optional cached ||row||^2 (bf16 / filter variants: replaces the in-kernel norm warps)
  const float* row_scale;  // fp16 shadow of an fp32 bank: 2^-e per row (multiplies the row's accumulators); else nullptr
  const float* qscale;     // fp16 shadow: 2^-f per query (from the query-prep kernel); else nullptr
  const uint32_t* need;    // optional [n_queries] flags: a CTA whose pass has none set exits at once (exact fallback)
  // filter variant only
  const float* qerr;       // [n_queries] relerr(dim) * ||query as used||: E = qerr * ||row|| (1 for normalised rows)
  const uint32_t* seed_u;  // optional [n_queries] score codes of a prefix sample's k-th best u (0 = none)
  const uint32_t* bank_max_sumsq;   // un-normalised scores: max ||row||^2 of the bank (float bits)
  int32_t k_true;          // the caller's top_k (lists keep kcap entries; T comes from entry k_true - 1)
  uint32_t* lostu;         // [grid, n_queries] out: score code bounding every row this CTA's full list lost (0 = none)
  uint32_t* gmax;          // [q_pad, gmax_stride] score code of each CTA's best upper bound per query (0 = none yet)
  int32_t gmax_stride;
  uint32_t* gthr;          // [n_queries] shared lower bounds of the k-th score (monotone score codes, 0 = none)
  // one-shot mode (short scans: every CTA owns exactly ONE tile, all CTAs co-resident): see oneshot_epilogue
  int32_t oneshot;
  int32_t emit;            // long scans: survivors appended to ocand instead of per-CTA lists (emit_epilogue)
  uint32_t* gbar;          // [2 * passes of the call] zeroed arrival counters: bests published / thresholds published
  uint32_t* ocount;        // [n_queries] zeroed: candidates emitted per query
  uint64_t* ocand;         // [n_queries, ocap] candidate keys (on the upper bound u), unordered
  int32_t ocap;
  IdMap idmap;
  long long* trace;        // optional [16][256] clock64 samples of CTA 0's roles (devqa_debug_tc_trace)
  int b_resident;          // filter, no pairs: the pass's whole query operand is loaded once and stays in shared memory
  CUtensorMap qmap;        // CTA pairs: the query operand as [rows of 128 bytes] (each CTA loads its half of a k-block)
};

// Role timeline tracing (tools/tc_trace.py) is compiled in only with -DDEVQA_TC_TRACE.
__device__ __forceinline__ void trace_at(const TcParams& p, int role, uint32_t idx) {
#ifdef DEVQA_TC_TRACE
  if (p.trace && blockIdx.x == 0 && idx < 256) p.trace[role * 256 + idx] = clock64();
#endif
}

struct SmemLayout {        // byte offsets from the 1024-aligned base
  uint32_t a_raw, bq, lists, thr, scratch, norms, bars, tmem_slot, total;
};
// kcap = 0: a list-free mode (emit / one-shot) -- no lists and no hit queue, only three words per query of scratch.
// b_kblocks > 0: the WHOLE query operand of the pass stays resident (b_kblocks k-blocks) instead of one k-block per stage.
__host__ __device__ inline SmemLayout tc_layout(int stages, int qp, int kcap, int b_parts = 2, bool pair = false, int b_kblocks = 0) {
  SmemLayout L;
  uint32_t off = 0;
  L.a_raw = off; off += uint32_t(stages) * kATileBytes;
  // [parts*qp rows x 128 B], 1024-aligned since qp % 16 == 0; a CTA pair holds half of the query rows each
  L.bq = off;    off += uint32_t(b_kblocks > 0 ? b_kblocks : stages) * uint32_t(b_parts * (pair ? qp / 2 : qp)) * 128;
  L.lists = off; off += uint32_t(qp) * kcap * 8;
  L.thr = off;   off += uint32_t(qp) * 20 + 16;                          // filter: eq / threshold / lost / service bound / query scale per query + counters
  L.scratch = off; off += kcap == 0 ? uint32_t(qp) * 12 + 16 : hit_cap(qp) * 8;   // hit queue / dense [16][128] score tile (MULTI: score tile [HQ][128], same size)
  L.norms = off; off += qp > 128 ? 0 : kNormRing * kTileRows * 4;         // (256 queries per pass: cached row norms only)
  off = (off + 7) & ~7u;
  L.bars = off;  off += (4 * kMaxStages + 2 * kMaxAccBufs) * 8;   // full_tma, full_split, empty, peer_full ([0]: resident query operand), tmem_full/empty
  L.tmem_slot = off; off += 16;
  L.total = off;
  return L;
}

// Epilogue phase B, slow path for one query whose list this warp owns exclusively.  Returns the new
// threshold (score of the k-th entry).  Deliberately out of line and loop-shaped (small, I-cache hot).
// Shared memory is the kernel's scarcest resource (its pipe is ~80% busy with TMA writes, split reads and
// UMMA operand fetches, so every dependent LDS/STS costs hundreds of cycles): for the common capacity 32
// the list is pulled into registers once (one key per lane), all inserts are shuffles, and it is written
// back once.
__device__ __noinline__ float epilogue_insert_query(uint64_t* list, int kcap, int top_k, const float* tile_row, float th,
                                                    uint32_t rowmask, int64_t row0, IdMap idmap, int lane) {
  const float4 sc = *reinterpret_cast<const float4*>(tile_row + lane * 4);
  const uint32_t bits = ((!(sc.x < th) ? 1u : 0u) | (!(sc.y < th) ? 2u : 0u) | (!(sc.z < th) ? 4u : 0u) | (!(sc.w < th) ? 8u : 0u)) & rowmask;
  uint64_t mylist = kcap == 32 ? list[lane] : 0ull;
  for (int t = 0; t < 4; ++t) {
    const bool on = (bits >> t) & 1u;
    unsigned m = __ballot_sync(kFull, on);
    if (!m) continue;
    const float st = t == 0 ? sc.x : (t == 1 ? sc.y : (t == 2 ? sc.z : sc.w));
    const uint64_t mine = on ? pack_key(st, uint32_t(idmap(row0 + t))) : kEmptyKey;
    if (kcap == 32) {
      if (__popc(m) >= 6) {                               // burst: sort the candidates once and merge
        const uint64_t sorted = warp_sort32_desc(mine, lane);
        mylist = warp_merge32(mylist, shfl_u64(sorted, 31 - lane), lane);
      } else {
        while (m) {
          const int src = __ffs(m) - 1;
          m &= m - 1;
          const uint64_t x = shfl_u64(mine, src);
          const int pos = __popc(__ballot_sync(kFull, mylist > x));
          const uint64_t up = shfl_u64(mylist, lane > 0 ? lane - 1 : 0);
          if (pos < top_k) mylist = lane > pos ? up : (lane == pos ? x : mylist);
        }
      }
    } else {
      while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        warp_list_insert(list, kcap, top_k, shfl_u64(mine, src), lane);
      }
    }
  }
  uint64_t kth;
  if (kcap == 32) {
    list[lane] = mylist;
    kth = shfl_u64(mylist, top_k - 1);
  } else {
    __syncwarp();
    kth = list[top_k - 1];
  }
  return kth == kEmptyKey ? -INFINITY : key_score(kth);
}

// ---- filter variant: epilogue ------------------------------------------------------------------------------
// Thread <-> TMEM lane <-> bank row.  Each thread walks its row's QP accumulators (16 per tcgen05.ld), forms the
// upper bound u = s~ + E and compares it with that query's threshold (shared memory, broadcast reads).  Nothing
// is stored for the overwhelming majority that fail; the rare survivors are pushed into a shared-memory hit queue
// and, after ONE 128-thread barrier per tile, inserted into the per-query lists by the warp that owns the query
// (q % 4).  The previous epilogue wrote every score to shared memory and re-read it (64 KB per tile at 128
// queries, on a shared-memory pipe that TMA and UMMA already keep ~90% busy): 93 cycles per query and tile
// measured against ~35 here.  A tile that overflows the queue (cold start: no thresholds yet) is redone in
// segments of 16 queries through a dense [16][128] score tile and the burst-capable list insertion of the
// other modes.
//
// Thresholds.  A row can be dropped when u < T for any lower bound T of the exact k-th best score.  Locally,
// T = (the list's k-th best u) - 2E.  Globally: every CTA publishes the score code of its best u per query
// (gmax); if k CTAs each hold a row with u >= x then k rows have s >= x - 2E, so T = (k-th largest of the
// CTAs' bests) - 2E.  Because the global top-k rows sit in (almost) distinct CTAs this is within a few rows of
// the ideal "k-th best of everything seen so far", whereas the best LOCAL bound only knows 1/148 of the rows
// (measured: 0.29 -> <0.01 list insertions per query and tile).  The otherwise idle warps 3-7 run this
// selection continuously (service warps) and hand the result to the epilogue through shared memory.
struct FiltState {            // per-query state in shared memory, [QP] each
  float* qs;                  // fp16 shadow: 2^-f of the query's power-of-two scaling (1 otherwise)
  float* eq;                  // tc_filter_relerr * ||query||
  float* thr;                 // admission threshold on u: max(lower bounds T of the exact k-th best, full list's last entry)
  float* lost;                // upper bound of every row a full list turned away while its last entry was binding
  float* thr_ext;             // T from the service warps (written by them only)
  uint32_t* qn;               // [2] hit counters (alternating per segment), [2] = scan finished
  float* hit_u;               // [hit_cap(QP)]   (dense mode: score tile [16][128], running into hit_id)
  uint32_t* hit_id;           // [hit_cap(QP)] query << 8 | row in tile
};

// After the list of query j changed: derive the local bound, keep the lost-rows bound, publish the CTA's best.
__device__ __forceinline__ void filt_after_insert(const TcParams& p, const FiltState& st, int j, int q0, int cta, float maxn,
                                                  uint64_t best, uint64_t ktrue, uint64_t kth) {
  float th = st.thr[j];
  // k-th best upper bound - 2E  <=  k-th best exact score of these rows  <=  that of the bank
  if (ktrue != kEmptyKey) {
    const float t = key_score(ktrue) - 2.f * st.eq[j] * maxn;
    if (t > th) th = t;
  }
  // A full list admits nothing below its last entry.  Rows it turns away while that entry is the binding
  // threshold (at or above T) are LOST, not proven irrelevant: remember their upper bound for refine.cu.
  if (kth != kEmptyKey) {
    const float l = key_score(kth);
    if (l >= th) { th = l; st.lost[j] = l; }
  }
  st.thr[j] = th;
  if (p.gmax) p.gmax[size_t(q0 + j) * p.gmax_stride + cta] = uint32_t(best >> 32);   // single writer per slot
}

// PARTS = 1: the filter variant (u = s~ + E).  PARTS = 3: the exact variants (3xTF32 / bf16x3; the score is the sum of
// three accumulator column groups, E = 0, so u IS the score and T the k-th best itself): same scan, queue and exchange.
template <int QP, int PARTS>
__device__ __forceinline__ void filt_epilogue(const TcParams& p, const FiltState st, uint64_t* lists, const float* norms,
                                              bool norm_warps, uint32_t tmem_base, uint64_t* tmem_full, uint64_t* tmem_empty,
                                              int tile0, int tile_step, int n_my, int boot, int q0, int nq_pass, float maxn,
                                              int e, int lane, uint32_t tmem_empty_remote = 0u) {
  // accumulator buffer a may be overwritten (pairs: the peer CTA tells the leader, who issues the MMAs)
  auto release_tmem = [&](int a) {
    if (tmem_empty_remote) mbar_arrive_cluster(tmem_empty_remote + uint32_t(a) * 8u);
    else mbar_arrive(&tmem_empty[a]);
  };
  const int row_in_tile = e * 32 + lane;
  const bool no_insert = (p.flags & kDbgNoInsert) != 0;
  const bool dot_mode = !(p.flags & DEVQA_NORMALIZE_BANK);          // E scales with ||row||
  uint32_t lt = 0, unit = 0, seg_idx = 0;
  int slow_calls = 0;
  uint32_t* const best_s = st.hit_id;                                // bootstrap tiles: best score code per query
  if (boot) {
    for (int j = e * 32 + lane; j < QP; j += 128) best_s[j] = 0u;
    epilogue_barrier();
  }
  // cached row norms: the NEXT tile's value is requested a whole tile ahead (a global load per tile at the point of
  // use cost ~800 cycles of every 7000-cycle tile period on the shadow path)
  const bool cached = !norm_warps && (!dot_mode || PARTS == 1);
  auto tile_of = [&](int it) { return tile0 + (it < n_my ? it : it - n_my) * tile_step; };
  auto ssq_of = [&](int it) {
    const int64_t g = int64_t(tile_of(it)) * kTileRows + row_in_tile;
    return (cached && it < n_my + boot && g < p.n_rows) ? __ldg(p.row_sumsq + g) : 1.f;
  };
  float ssq_next = ssq_of(0);
  for (int it = 0; it < n_my + boot; ++it, ++lt, ++unit) {
    const int tile = tile_of(it);
    const int a = unit & 1;
    const float ssq_cached = ssq_next;
    ssq_next = ssq_of(it + 1);
    mbar_wait(&tmem_full[a], (unit >> 1) & 1);
    tc_fence_after();
    if (e == 0 && lane == 0) trace_at(p, 6, unit);
    const int64_t grow = int64_t(tile) * kTileRows + row_in_tile;
    const bool row_live = grow < p.n_rows && !no_insert;
    float inv = 1.f, rowfac = 1.f;
    if (!dot_mode || PARTS == 1) {
      const float ssq = norm_warps ? norms[(lt % kNormRing) * kTileRows + row_in_tile] : ssq_cached;
      const float nrm = sqrtf(ssq);
      if (dot_mode) rowfac = nrm;
      else inv = 1.0f / ((p.flags & DEVQA_NOEPS) ? nrm : fmaxf(nrm, kNormEps));
    }
    const uint32_t taddr = tmem_base + uint32_t(a) * (PARTS * QP) + (uint32_t(e * 32) << 16);
    // 16 accumulators of this row -> 16 upper bounds (filter) / scores (exact variants)
    auto load_u = [&](int c, float (&u)[16]) {
      uint32_t d[16];
      tmem_ld<16>(taddr + uint32_t(c), d);
      if constexpr (PARTS == 3) {
        uint32_t d2[16], d3[16];
        tmem_ld<16>(taddr + uint32_t(QP + c), d2);
        tmem_ld<16>(taddr + uint32_t(2 * QP + c), d3);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i)
          u[i] = ((__uint_as_float(d2[i]) + __uint_as_float(d3[i])) + __uint_as_float(d[i])) * inv;
      } else {
        tmem_ld_wait();
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          const float4 eq4 = *reinterpret_cast<const float4*>(st.eq + c + 4 * i4);
          u[4 * i4 + 0] = fmaf(__uint_as_float(d[4 * i4 + 0]), inv, eq4.x * rowfac);
          u[4 * i4 + 1] = fmaf(__uint_as_float(d[4 * i4 + 1]), inv, eq4.y * rowfac);
          u[4 * i4 + 2] = fmaf(__uint_as_float(d[4 * i4 + 2]), inv, eq4.z * rowfac);
          u[4 * i4 + 3] = fmaf(__uint_as_float(d[4 * i4 + 3]), inv, eq4.w * rowfac);
        }
      }
    };
    if (it < boot) {
      // ---- bootstrap visit: only each query's best upper bound in this tile, for the threshold exchange ----
#pragma unroll 1
      for (int c = 0; c < QP; c += 16) {
        float u[16];
        load_u(c, u);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const uint32_t m = __reduce_max_sync(kFull, row_live ? score_code(u[i]) : 0u);
          if (lane == 0) atomicMax(best_s + c + i, m);
        }
      }
      tc_fence_before();
      release_tmem(a);
      epilogue_barrier();
      for (int j = e * 32 + lane; j < nq_pass; j += 128)
        p.gmax[size_t(q0 + j) * p.gmax_stride + tile0] = best_s[j];
      for (int j = 4 * lane + e; j < nq_pass; j += 128) {              // thresholds the service warps have derived so far
        {
          if (it == boot - 1) {                                        // last visit: give them a bounded moment to arrive --
            const long long t0 = clock64();                            // a first ordinary tile without thresholds costs ~30 tiles
            while (*reinterpret_cast<volatile float*>(st.thr_ext + j) == -INFINITY && clock64() - t0 < 100000) __nanosleep(200);
          }
          st.thr[j] = fmaxf(st.thr[j], *reinterpret_cast<volatile float*>(st.thr_ext + j));
        }
      }
      epilogue_barrier();                                              // (also: best_s words go back to the hit queue)
      continue;
    }
    // dense = 16 queries at a time through a score tile: after a queue overflow, and for a first tile without
    // thresholds (short scans: every row is a candidate)
    bool dense = it == 0 && boot == 0;
    int seg = dense ? 16 : QP;
    for (int jb = 0; jb < QP; jb += seg, ++seg_idx) {
      uint32_t* const qn = st.qn + (seg_idx & 1);
      // ---- scan ----
      // 16 scores of this row against their thresholds; the few survivors go to the hit queue (or, dense: all 16 to the tile)
      auto consume = [&](int c, const float (&u)[16]) {
        uint32_t mask = 0;
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          const float4 th4 = *reinterpret_cast<const float4*>(st.thr + c + 4 * i4);
          mask |= (!(u[4 * i4 + 0] < th4.x) ? 1u : 0u) << (4 * i4 + 0);   // NaN passes, like torch.topk ranks it first
          mask |= (!(u[4 * i4 + 1] < th4.y) ? 1u : 0u) << (4 * i4 + 1);
          mask |= (!(u[4 * i4 + 2] < th4.z) ? 1u : 0u) << (4 * i4 + 2);
          mask |= (!(u[4 * i4 + 3] < th4.w) ? 1u : 0u) << (4 * i4 + 3);
        }
        if (!row_live) mask = 0u;
        if (dense) {                                                    // the segment's scores as a [16][128] tile
#pragma unroll
          for (int i = 0; i < 16; ++i) st.hit_u[i * kTileRows + row_in_tile] = u[i];
        } else if (__any_sync(kFull, mask != 0u)) {                     // rare
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if ((mask >> i) & 1u) {
              const uint32_t idx = atomicAdd(qn, 1u);
              if (idx < uint32_t(hit_cap(QP))) { st.hit_u[idx] = u[i]; st.hit_id[idx] = (uint32_t(c + i) << 8) | uint32_t(row_in_tile); }
            }
          }
        }
      };
      if constexpr (PARTS == 1) {
        // filter variants: the tensor-memory load of the NEXT 16 accumulators is in flight while these 16 are compared
        // (tcgen05.wait::ld waits for everything outstanding, so: wait, issue the next, then work on the current)
        auto to_u = [&](int c, const uint32_t (&d)[16], float (&u)[16]) {
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            const float4 eq4 = *reinterpret_cast<const float4*>(st.eq + c + 4 * i4);
            u[4 * i4 + 0] = fmaf(__uint_as_float(d[4 * i4 + 0]), inv, eq4.x * rowfac);
            u[4 * i4 + 1] = fmaf(__uint_as_float(d[4 * i4 + 1]), inv, eq4.y * rowfac);
            u[4 * i4 + 2] = fmaf(__uint_as_float(d[4 * i4 + 2]), inv, eq4.z * rowfac);
            u[4 * i4 + 3] = fmaf(__uint_as_float(d[4 * i4 + 3]), inv, eq4.w * rowfac);
          }
        };
        uint32_t da[16], db[16];
        const int c_end = jb + seg;
        tmem_ld<16>(taddr + uint32_t(jb), da);
#pragma unroll 1
        for (int c = jb; c < c_end; c += 32) {
          float u[16];
          tmem_ld_wait();
          if (c + 16 < c_end) tmem_ld<16>(taddr + uint32_t(c + 16), db);
          to_u(c, da, u);
          consume(c, u);
          if (c + 16 < c_end) {
            tmem_ld_wait();
            if (c + 32 < c_end) tmem_ld<16>(taddr + uint32_t(c + 32), da);
            to_u(c + 16, db, u);
            consume(c + 16, u);
          }
        }
      } else {
#pragma unroll 1
        for (int c = jb; c < jb + seg; c += 16) {
          float u[16];
          load_u(c, u);
          consume(c, u);
        }
      }
      const bool last_seg = jb + seg >= QP;
      tc_fence_before();
      epilogue_barrier();                                               // every hit of the segment is visible
      const uint32_t n = dense ? 0u : *qn;
      if (e == 0 && lane == 0) st.qn[(seg_idx & 1) ^ 1] = 0u;           // the other counter: idle until the next scan
      // Many hits (cold lists, short scans): one-by-one insertion costs ~1000 cycles each per warp, the dense path a
      // fixed ~9k cycles per 16 queries -- redo the tile densely beyond the break-even (and always on overflow).
      if (n > uint32_t(2 * QP + 32 < hit_cap(QP) ? 2 * QP + 32 : hit_cap(QP))) {   // (cold start): redo the tile densely
        dense = true;
        seg = 16;
        jb = -16;
        epilogue_barrier();
        continue;
      }
      if (last_seg) release_tmem(a);                                    // this thread's TMEM reads of the buffer are done
      if (dense) {
        // ---- dense segment: warp e owns queries jb + e, jb + 4 + e, ... ----
        const int64_t row0 = int64_t(tile) * kTileRows + lane * 4;
        uint32_t rowmask = 0;
#pragma unroll
        for (int t = 0; t < 4; ++t) rowmask |= (row0 + t < p.n_rows ? 1u : 0u) << t;
        if (no_insert) rowmask = 0;
#pragma unroll 1
        for (int i = e; i < 16; i += 4) {
          const int j = jb + i;
          if (j >= nq_pass) break;
          const float th = st.thr[j];
          const float4 sc = *reinterpret_cast<const float4*>(st.hit_u + i * kTileRows + lane * 4);
          const bool hit = (!(sc.x < th) && (rowmask & 1u)) || (!(sc.y < th) && (rowmask & 2u)) ||
                           (!(sc.z < th) && (rowmask & 4u)) || (!(sc.w < th) && (rowmask & 8u));
          if (!__any_sync(kFull, hit)) continue;
          ++slow_calls;
          uint64_t* list = lists + size_t(j) * p.kcap;
          epilogue_insert_query(list, p.kcap, p.top_k, st.hit_u + i * kTileRows, th, rowmask, row0, p.idmap, lane);
          __syncwarp();
          if (lane == 0) filt_after_insert(p, st, j, q0, tile0, maxn, list[0], list[p.k_true - 1], list[p.top_k - 1]);
          __syncwarp();
        }
      } else {
        // ---- drain: warp e owns the queries with q % 4 == e ----
        // The queue is read 32 hits at a time, ONE load per lane; the warp then walks its own hits through shuffles
        // and has every shared-memory word an insertion needs (the list, the threshold, the error scale) in flight
        // together.  The previous form chained six dependent shared-memory round trips per hit on a pipe that TMA and
        // UMMA keep ~90% busy -- ~1000 cycles per insertion, and with the 2-byte shadow (tile period 7 k cycles instead
        // of 14 k) the epilogue, not HBM, set the pace.
        for (uint32_t h0 = 0; h0 < n; h0 += 32) {
          const uint32_t h = h0 + uint32_t(lane);
          uint32_t my_id = 0u;
          float my_u = 0.f;
          if (h < n) { my_id = st.hit_id[h]; my_u = st.hit_u[h]; }
          unsigned todo = __ballot_sync(kFull, h < n && (int(my_id >> 8) & 3) == e);
          while (todo) {                                                // warp-uniform
            const int srcl = __ffs(todo) - 1;
            todo &= todo - 1;
            const uint32_t id = __shfl_sync(kFull, my_id, srcl);
            const float uu = __shfl_sync(kFull, my_u, srcl);
            const int j = int(id >> 8);
            uint64_t* list = lists + size_t(j) * p.kcap;
            const float th_now = st.thr[j];                             // (three independent loads)
            const float eq_j = st.eq[j];
            uint64_t mine = p.kcap == 32 ? list[lane] : 0ull;
            if (uu < th_now) continue;                                  // the threshold has risen since the push
            ++slow_calls;
            const uint64_t key = pack_key(uu, uint32_t(p.idmap(int64_t(tile) * kTileRows + int64_t(id & 0xffu))));
            uint64_t best, kth, ktrue;
            if (p.kcap == 32) {
              const int pos = __popc(__ballot_sync(kFull, mine > key));
              const uint64_t up = shfl_u64(mine, lane > 0 ? lane - 1 : 0);
              if (pos < p.top_k) mine = lane > pos ? up : (lane == pos ? key : mine);
              list[lane] = mine;
              best = shfl_u64(mine, 0);
              kth = shfl_u64(mine, p.top_k - 1);
              ktrue = shfl_u64(mine, p.k_true - 1);
            } else {
              warp_list_insert(list, p.kcap, p.top_k, key, lane);
              __syncwarp();
              best = list[0];
              kth = list[p.top_k - 1];
              ktrue = list[p.k_true - 1];
            }
            if (lane == 0) {                                            // filt_after_insert, on values already in registers
              float th = th_now;
              if (ktrue != kEmptyKey) {
                const float t = key_score(ktrue) - 2.f * eq_j * maxn;
                if (t > th) th = t;
              }
              if (kth != kEmptyKey) {
                const float l = key_score(kth);
                if (l >= th) { th = l; st.lost[j] = l; }
              }
              st.thr[j] = th;
              if (p.gmax) p.gmax[size_t(q0 + j) * p.gmax_stride + tile0] = uint32_t(best >> 32);   // single writer per slot
            }
            __syncwarp();
          }
        }
      }
      // ---- pick up the service warps' bounds (own queries only: nobody else writes thr) ----
      if (last_seg)
        for (int j = 4 * lane + e; j < nq_pass; j += 128) st.thr[j] = fmaxf(st.thr[j], st.thr_ext[j]);
      epilogue_barrier();                                               // thresholds visible, queue reusable
    }
    if (e == 0 && lane == 0) trace_at(p, 7, unit);
#ifdef DEVQA_TC_TRACE
    if (e == 0 && lane == 0 && p.trace && blockIdx.x == 0 && unit < 256) { p.trace[8 * 256 + unit] = slow_calls; slow_calls = 0; }
#endif
  }
  if (e == 0 && lane == 0) *reinterpret_cast<volatile uint32_t*>(st.qn + 2) = 1u;   // service warps: stop
  // what this CTA's lists lost, for the completeness proof
  if (p.lostu)
    for (int j = e * 32 + lane; j < nq_pass; j += 128) {
      const float l = st.lost[j];
      p.lostu[size_t(tile0) * p.n_queries + q0 + j] = l == -INFINITY ? 0u : score_code(l);
    }
}

// Grid-wide arrival counters in global memory (all CTAs of the launch are co-resident: one per SM, grid <= SM count).
__device__ __forceinline__ void grid_arrive(uint32_t* ctr) {
  __threadfence();
  atomicAdd(ctr, 1u);
}
// returns false when it gave up (the caller then carries on without what it waited for; results stay correct, the
// queries concerned overflow their candidate arrays and are redone exactly)
__device__ __forceinline__ bool grid_wait(const uint32_t* ctr, uint32_t want, long long max_cycles = 2000000000ll) {
  const long long t0 = clock64();
  uint32_t v;
  do {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    if (v >= want) return true;
    __nanosleep(100);
  } while (clock64() - t0 < max_cycles);
  return false;
}

// One round of the shared-out threshold selection: CTA `cta` derives T for the queries j = cta, cta + n_ctas, ... of
// its pass from every CTA's best upper bound (gmax) and publishes it in gthr (atomicMax on the score code).
__device__ __forceinline__ void service_select_round(const TcParams& p, const FiltState& st, int q0, int nq_pass, int n_ctas,
                                                     int cta, float maxn, int lane, int k) {
  constexpr int NQ = 4;                                               // queries selected side by side (hides the redux latency)
  for (int jg = cta; jg < nq_pass; jg += n_ctas * NQ) {
    uint32_t v[NQ][5];
#pragma unroll
    for (int t = 0; t < NQ; ++t) {
      const int j = jg + t * n_ctas < nq_pass ? jg + t * n_ctas : jg;
      const uint32_t* src = p.gmax + size_t(q0 + j) * p.gmax_stride;
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        const int c = lane + 32 * i;
        v[t][i] = 0u;
        if (c < n_ctas) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v[t][i]) : "l"(src + c));
      }
    }
    // per query: the largest x (to 20 bits) with at least k values >= x, a lower bound of the k-th largest
    uint32_t x[NQ];
#pragma unroll
    for (int t = 0; t < NQ; ++t) x[t] = 0u;
    const int n_live = min(NQ, (nq_pass - jg + n_ctas - 1) / n_ctas);   // slots of this group that hold a query (warp-uniform)
#pragma unroll 1
    for (int bit = 31; bit >= 12; --bit) {
#pragma unroll
      for (int t = 0; t < NQ; ++t) {
        if (t >= n_live) break;
        const uint32_t cand = x[t] | (1u << bit);
        int cnt = 0;
#pragma unroll
        for (int i = 0; i < 5; ++i) cnt += v[t][i] >= cand ? 1 : 0;
        cnt = __reduce_add_sync(kFull, cnt);
        if (cnt >= k) x[t] = cand;
      }
    }
#pragma unroll
    for (int t = 0; t < NQ; ++t) {
      const int j = jg + t * n_ctas;
      if (lane == 0 && j < nq_pass && x[t] != 0u) {
        const float tt = code_score(x[t]) - 2.f * st.eq[j] * maxn;   // NaN (k CTAs hold NaN rows) maps to the top code:
        if (tt == tt) atomicMax(p.gthr + q0 + j, score_code(tt));   // not a usable bound, skip it
      }
    }
  }
}

// Service warp (warp 3).  The selection is shared out: CTA c derives T for the queries j = c, c + n_ctas, ... from
// all CTAs' bests and publishes it (gthr, one word per query); every CTA then picks up all of its pass's
// thresholds with one coalesced read.  (Each CTA selecting for every query cost 148x the work and 16-32 us per
// round at 128-256 queries; this way a round is a few microseconds.)
__device__ __forceinline__ void filt_service(const TcParams& p, const FiltState st, int q0, int nq_pass, int n_ctas, int cta,
                                             float maxn, int lane) {
  volatile uint32_t* const done = st.qn + 2;
  const int k = p.k_true;
  if (n_ctas < k || !p.gmax) return;                                  // the bound needs k distinct CTAs
  volatile float* const thr_ext = st.thr_ext;
  if (p.emit && p.gbar) {
    // Emit mode, first exchange.  The epilogue's first tile WAITS for thresholds, with the MMA pipeline stalled behind
    // it: polling rounds (select ~3 k cycles + sleep, once here and once in the CTA that owns the query) made that
    // wait ~28 k cycles (traced).  Instead, as in the one-shot mode: wait until every CTA has published its first-tile
    // bests (arrival counter), select, publish, wait until every CTA has (second counter), pick the thresholds up.
    // Bounded: if the grid is not co-resident the counters never fill, and the polling loop below takes over.
    const int pg = p.q0 / p.qp + int(blockIdx.x) % p.passes;
    uint32_t ok = 0u;
    if (lane == 0) ok = grid_wait(p.gbar + 2 * pg, uint32_t(n_ctas), 400000) ? 1u : 0u;
    ok = __shfl_sync(kFull, ok, 0);
    if (lane == 0) trace_at(p, 12, 0);
    if (ok) {
      service_select_round(p, st, q0, nq_pass, n_ctas, cta, maxn, lane, k);
      __syncwarp();
      if (lane == 0) trace_at(p, 13, 0);
      if (lane == 0) { grid_arrive(p.gbar + 2 * pg + 1); grid_wait(p.gbar + 2 * pg + 1, uint32_t(n_ctas), 400000); }
      __syncwarp();
      if (lane == 0) trace_at(p, 14, 0);
      for (int j = lane; j < nq_pass; j += 32) {
        uint32_t code;
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(code) : "l"(p.gthr + q0 + j));
        if (code != 0u) thr_ext[j] = code_score(code);
      }
    }
  }
  // The k-th best of everything seen grows like log(rows): refresh eagerly at first, then back off (a spinning
  // warp costs issue slots and, on a power-capped part, clock).
  unsigned pause_ns = 500;
  while (!*done) {
    service_select_round(p, st, q0, nq_pass, n_ctas, cta, maxn, lane, k);
    for (int j = lane; j < nq_pass; j += 32) {
      uint32_t code;
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(code) : "l"(p.gthr + q0 + j));
      if (code != 0u) {
        const float tt = code_score(code);
        if (tt > thr_ext[j]) thr_ext[j] = tt;
      }
    }
    for (unsigned slept = 0; slept < pause_ns && !*done; slept += 1000) __nanosleep(pause_ns < 1000 ? pause_ns : 1000);
    // back off only once thresholds exist: the epilogue's first tile WAITS for them (a service warp that had already
    // grown its pause to 8-16 us while the first tile was still being multiplied cost ~20 us per scan)
    bool have = true;
    for (int j = lane; j < nq_pass; j += 32) have = have && thr_ext[j] != -INFINITY;
    if (__all_sync(kFull, have) && pause_ns < 64000) pause_ns *= 2;
  }
}

// ---- emit mode: long scans on the filter path without per-CTA lists ----------------------------------------------------
// The list epilogue inserts every row that beats the running threshold into a sorted per-CTA list (~900 cycles per
// insertion on a shared-memory pipe that TMA and UMMA keep busy).  The number of such rows is set by the filter's error
// band: ~200 per query and scan with the TF32 filter, ~1500 with a bf16 shadow (the first version; band 2E = 0.43 sigma of the score
// distribution) -- there the epilogue, not HBM, set the tile period (traced: 7.1 k cycles of epilogue per 7.9 k-cycle
// tile; the first eight tiles 17 k each).  But nothing downstream needs the lists SORTED: refine.cu only wants every
// row with u >= T.  So, like the one-shot mode's pass B, a thread that finds a survivor appends it to a flat per-query
// candidate array in global memory (one atomicAdd + one store, no queue, no drain, no CTA barrier per tile) and
// raises the CTA's best-u word for the threshold exchange.  Thresholds come only from the exchange (service warp); the
// first tile's accumulators stay in tensor memory while the CTA waits for the first thresholds and are then scanned
// like any other tile (no second visit).  Nothing is ever pushed out of a list, so there are no lost-row bounds;
// a candidate array that overflows flags its query for the exact fallback (refine.cu).
template <int QP>
__device__ __forceinline__ void emit_epilogue(const TcParams& p, const FiltState st, const float* norms, bool norm_warps,
                                              uint32_t tmem_base, uint64_t* tmem_full, uint64_t* tmem_empty, int tile0,
                                              int tile_step, int n_my, int q0, int nq_pass, int e, int lane, int n_acc,
                                              uint32_t tmem_empty_remote, int g, int ng) {
  // ng = 2: a second set of four warps (g = 1: warps 4-7, idle when row norms are cached) takes the upper half of the
  // accumulator columns.  A warp reaches only the 32 tensor-memory lanes of its quarter, so more warps can only split
  // COLUMNS -- which is what a 256-query pass needs: one thread per row compared 256 scores per tile, 10.8 k cycles
  // against 6.1 k of tensor-core time.
  const int nthr = 128 * ng;
  const int c_lo = g * (QP / ng), c_hi = c_lo + QP / ng;
  auto ebar = [&]() { if (ng == 1) epilogue_barrier(); else epilogue_barrier_n(nthr); };
  auto release_tmem = [&](int a) {
    if (tmem_empty_remote) mbar_arrive_cluster(tmem_empty_remote + uint32_t(a) * 8u);
    else mbar_arrive(&tmem_empty[a]);
  };
  const int row_in_tile = e * 32 + lane;
  const int tid = g * 128 + e * 32 + lane;
  const bool dot_mode = !(p.flags & DEVQA_NORMALIZE_BANK);
  uint32_t* const best_s = st.hit_id;                                  // [QP] best upper bound this CTA has emitted, per query
  uint32_t* const pub_s = st.hit_id + QP;                              // [QP] what gmax holds
  uint32_t* const cnt_s = st.hit_id + 2 * QP;                          // [QP] candidates this CTA has emitted, per query
  for (int j = tid; j < QP; j += nthr) { best_s[j] = 0u; pub_s[j] = 0u; cnt_s[j] = 0u; }
  // this CTA's own region of the candidate array: [n_queries][ocap] -- private counters in shared memory, plain global
  // stores, nothing to wait for (a global atomicAdd per survivor stalled the finding thread for a full round trip)
  uint64_t* const my_cand = p.ocand + size_t(tile0) * p.n_queries * p.ocap;
  ebar();
  auto ssq_of = [&](int it) {
    const int64_t g = int64_t(tile0 + it * tile_step) * kTileRows + row_in_tile;
    return (!norm_warps && it < n_my && g < p.n_rows) ? __ldg(p.row_sumsq + g) : 1.f;
  };
  auto rs_of = [&](int it) {                                           // fp16 shadow: the row's 2^-e
    const int64_t g = int64_t(tile0 + it * tile_step) * kTileRows + row_in_tile;
    return (p.row_scale && it < n_my && g < p.n_rows) ? __ldg(p.row_scale + g) : 1.f;
  };
  const bool scaled = p.row_scale != nullptr;
  float ssq_next = ssq_of(0), rs_next = rs_of(0);
  for (int it = 0; it < n_my; ++it) {
    const int tile = tile0 + it * tile_step;
    const int a = it % n_acc;
    const float ssq_cached = ssq_next, rs_cached = rs_next;
    ssq_next = ssq_of(it + 1);
    rs_next = rs_of(it + 1);
    mbar_wait(&tmem_full[a], uint32_t(it / n_acc) & 1u);
    tc_fence_after();
    if (tid == 0) trace_at(p, 6, it);
    const int64_t grow = int64_t(tile) * kTileRows + row_in_tile;
    const bool row_live = grow < p.n_rows && !(p.flags & kDbgNoInsert);
    float inv = 1.f, rowfac = 1.f;
    {
      const float ssq = norm_warps ? norms[(it % kNormRing) * kTileRows + row_in_tile] : ssq_cached;
      const float nrm = sqrtf(ssq);
      if (dot_mode) rowfac = nrm;
      else inv = 1.0f / ((p.flags & DEVQA_NOEPS) ? nrm : fmaxf(nrm, kNormEps));
    }
    const uint32_t taddr = tmem_base + uint32_t(a) * QP + (uint32_t(e * 32) << 16);
    // approximate score = accumulator * (2^-e of the row) * (2^-f of the query) * 1/||row||: the two power-of-two
    // factors are exact, and they are applied BEFORE inv so that no intermediate leaves the float range
    auto to_u = [&](int c, const uint32_t (&d)[16], float (&u)[16]) {
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4) {
        const float4 eq4 = *reinterpret_cast<const float4*>(st.eq + c + 4 * i4);
        float4 f = make_float4(1.f, 1.f, 1.f, 1.f);
        if (scaled) {
          f = *reinterpret_cast<const float4*>(st.qs + c + 4 * i4);
          f.x *= rs_cached; f.y *= rs_cached; f.z *= rs_cached; f.w *= rs_cached;
        }
        u[4 * i4 + 0] = fmaf(__uint_as_float(d[4 * i4 + 0]) * f.x, inv, eq4.x * rowfac);
        u[4 * i4 + 1] = fmaf(__uint_as_float(d[4 * i4 + 1]) * f.y, inv, eq4.y * rowfac);
        u[4 * i4 + 2] = fmaf(__uint_as_float(d[4 * i4 + 2]) * f.z, inv, eq4.z * rowfac);
        u[4 * i4 + 3] = fmaf(__uint_as_float(d[4 * i4 + 3]) * f.w, inv, eq4.w * rowfac);
      }
    };
    if (it == 0) {
      // ---- first tile, pass A: each query's best upper bound -> exchange; then a bounded wait for the thresholds ----
#pragma unroll 1
      for (int c = c_lo; c < c_hi; c += 16) {
        uint32_t d[16];
        float u[16];
        tmem_ld<16>(taddr + uint32_t(c), d);
        tmem_ld_wait();
        to_u(c, d, u);
        // 16 warp-wide maxima back to back (each result is in every lane), then lane i posts the i-th: ONE shared-memory
        // atomic per 16 queries.  (A reduce + lane-0 atomic per query serialised to ~280 cycles each: 18 k cycles of
        // the first tile, with the MMA pipeline stalled behind it.)
        uint32_t mine = 0u;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const uint32_t m = __reduce_max_sync(kFull, row_live ? score_code(u[i]) : 0u);
          mine = lane == i ? m : mine;
        }
        if (lane < 16) atomicMax(best_s + c + lane, mine);
      }
      ebar();
      if (tid == 0) trace_at(p, 9, 0);
      for (int j = tid; j < nq_pass; j += nthr) {
        const uint32_t b = best_s[j];
        pub_s[j] = b;
        p.gmax[size_t(q0 + j) * p.gmax_stride + tile0] = b;
      }
      if (p.gbar) {                                                    // tell the service warps: this CTA's bests are in gmax
        __threadfence();
        ebar();
        if (tid == 0) atomicAdd(p.gbar + 2 * (p.q0 / p.qp + int(blockIdx.x) % p.passes), 1u);
      }
      if (tid == 0) trace_at(p, 10, 0);
      for (int j = tid; j < nq_pass; j += nthr) {
        const long long t0 = clock64();                                // (no thresholds: every row would be a candidate)
        while (*reinterpret_cast<volatile float*>(st.thr_ext + j) == -INFINITY && clock64() - t0 < 400000) __nanosleep(200);
        st.thr[j] = fmaxf(st.thr[j], *reinterpret_cast<volatile float*>(st.thr_ext + j));
      }
      ebar();
      if (tid == 0) trace_at(p, 11, 0);
    }
    // ---- scan: 16 accumulators at a time, the next tensor-memory load in flight while these are compared ----
    auto consume = [&](int c, const float (&u)[16]) {
      uint32_t mask = 0;
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4) {
        const float4 th4 = *reinterpret_cast<const float4*>(st.thr + c + 4 * i4);
        mask |= (!(u[4 * i4 + 0] < th4.x) ? 1u : 0u) << (4 * i4 + 0);     // NaN passes, like torch.topk ranks it first
        mask |= (!(u[4 * i4 + 1] < th4.y) ? 1u : 0u) << (4 * i4 + 1);
        mask |= (!(u[4 * i4 + 2] < th4.z) ? 1u : 0u) << (4 * i4 + 2);
        mask |= (!(u[4 * i4 + 3] < th4.w) ? 1u : 0u) << (4 * i4 + 3);
      }
      if (!row_live) mask = 0u;
      // (a NaN passes every comparison, also the +inf threshold of a padding query: those have no candidate region)
      if (c + 16 > nq_pass) mask &= c < nq_pass ? (1u << (nq_pass - c)) - 1u : 0u;
      if (mask != 0u) {                                                // rare
        const uint32_t id = uint32_t(p.idmap(grow));
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if ((mask >> i) & 1u) {
            atomicMax(best_s + c + i, score_code(u[i]));
            const uint32_t idx = atomicAdd(cnt_s + c + i, 1u);
            if (idx < uint32_t(p.ocap)) my_cand[size_t(q0 + c + i) * p.ocap + idx] = pack_key(u[i], id);
          }
        }
      }
    };
    {
      uint32_t da[16], db[16];
      tmem_ld<16>(taddr + uint32_t(c_lo), da);
#pragma unroll 1
      for (int c = c_lo; c < c_hi; c += 32) {
        float u[16];
        tmem_ld_wait();
        if (c + 16 < c_hi) tmem_ld<16>(taddr + uint32_t(c + 16), db);
        to_u(c, da, u);
        consume(c, u);
        if (c + 16 < c_hi) {
          tmem_ld_wait();
          if (c + 32 < c_hi) tmem_ld<16>(taddr + uint32_t(c + 32), da);
          to_u(c + 16, db, u);
          consume(c + 16, u);
        }
      }
    }
    tc_fence_before();
    release_tmem(a);                                                   // this thread's reads of the buffer are done
    // ---- thresholds from the service warp; this CTA's bests to the exchange.  No barrier: thr only ever rises, and a
    // reader that still sees the old word merely emits a candidate more ----
    for (int j = tid; j < nq_pass; j += nthr) {
      st.thr[j] = fmaxf(st.thr[j], st.thr_ext[j]);
      const uint32_t b = best_s[j];
      if (b != pub_s[j]) { pub_s[j] = b; p.gmax[size_t(q0 + j) * p.gmax_stride + tile0] = b; }
    }
    if (tid == 0) trace_at(p, 7, it);
  }
  ebar();
  if (tid == 0) *reinterpret_cast<volatile uint32_t*>(st.qn + 2) = 1u;   // service warp: stop
  for (int j = tid; j < nq_pass; j += nthr) {
    const uint32_t cnt = cnt_s[j];
    p.ocount[size_t(tile0) * p.n_queries + q0 + j] = cnt < uint32_t(p.ocap) ? cnt : uint32_t(p.ocap);
    // nothing is ever lost -- unless this CTA's region overflowed: then "anything may be missing" (NaN code: refine
    // flags the query for the exact fallback)
    p.lostu[size_t(tile0) * p.n_queries + q0 + j] = cnt > uint32_t(p.ocap) ? 0xffffffffu : 0u;
  }
}

// ---- one-shot mode: short scans (the whole bank is at most one tile per SM) -------------------------------------------
// C2's shape -- 8174 rows x 4096 columns, 256 queries -- is 64 tiles: every CTA owns exactly ONE tile and all CTAs are
// co-resident.  The persistent kernel's epilogue is built for long scans (thresholds learnt over many tiles, per-CTA
// lists, ~1000 cycles per list insertion): on a CTA's first tile there are no thresholds, so every row is a candidate
// and list building cost 75 us of a 134 us scan.  Here instead, with the tile's accumulators staying in tensor memory:
//   pass A   each query's best upper bound u in this tile (one redux.max per query and warp) -> gmax[q][tile];
//   barrier  all CTAs of the pass have published (arrival counter in global memory, spin -- co-residency is guaranteed
//            by the cooperative launch);
//   select   CTA c derives T = (k-th largest of the tiles' bests) - 2E for queries c, c + n_ctas, ... (service warp) and
//            publishes it; second barrier;
//   pass B   the SAME accumulators are scanned again against T: the few rows with u >= T (about k + 10 per query over the
//            whole bank) are appended to a global candidate array by atomic counter -- no lists, nothing is ever "lost".
// refine.cu then re-scores the candidates exactly (refine_lists_kernel, direct mode).  A query whose candidates overflow
// the array (no usable T: NaN, fewer than k tiles) is flagged there and redone by the exact pipeline.
template <int QP>
__device__ __forceinline__ void oneshot_epilogue(const TcParams& p, const FiltState st, const float* norms, bool norm_warps,
                                                 uint32_t tmem_base, uint64_t* tmem_full, int tile, int n_ctas, int pass_global,
                                                 int q0, int nq_pass, int e, int lane) {
  const int row_in_tile = e * 32 + lane;
  const bool dot_mode = !(p.flags & DEVQA_NORMALIZE_BANK);
  mbar_wait(&tmem_full[0], 0);
  tc_fence_after();
  const int64_t grow = int64_t(tile) * kTileRows + row_in_tile;
  const bool row_live = grow < p.n_rows && !(p.flags & kDbgNoInsert);
  float inv = 1.f, rowfac = 1.f;
  {
    const float ssq = norm_warps ? norms[row_in_tile] : (grow < p.n_rows ? p.row_sumsq[grow] : 1.f);
    const float nrm = sqrtf(ssq);
    if (dot_mode) rowfac = nrm;
    else inv = 1.0f / ((p.flags & DEVQA_NOEPS) ? nrm : fmaxf(nrm, kNormEps));
  }
  const uint32_t taddr = tmem_base + (uint32_t(e * 32) << 16);
  const bool scaled = p.row_scale != nullptr;                          // fp16 shadow: power-of-two scales of row and query
  const float rs = (scaled && grow < p.n_rows) ? __ldg(p.row_scale + grow) : 1.f;
  auto load_u = [&](int c, float (&u)[16]) {
    uint32_t d[16];
    tmem_ld<16>(taddr + uint32_t(c), d);
    tmem_ld_wait();
#pragma unroll
    for (int i4 = 0; i4 < 4; ++i4) {
      const float4 eq4 = *reinterpret_cast<const float4*>(st.eq + c + 4 * i4);
      float4 f = make_float4(1.f, 1.f, 1.f, 1.f);
      if (scaled) {
        f = *reinterpret_cast<const float4*>(st.qs + c + 4 * i4);
        f.x *= rs; f.y *= rs; f.z *= rs; f.w *= rs;
      }
      u[4 * i4 + 0] = fmaf(__uint_as_float(d[4 * i4 + 0]) * f.x, inv, eq4.x * rowfac);
      u[4 * i4 + 1] = fmaf(__uint_as_float(d[4 * i4 + 1]) * f.y, inv, eq4.y * rowfac);
      u[4 * i4 + 2] = fmaf(__uint_as_float(d[4 * i4 + 2]) * f.z, inv, eq4.z * rowfac);
      u[4 * i4 + 3] = fmaf(__uint_as_float(d[4 * i4 + 3]) * f.w, inv, eq4.w * rowfac);
    }
  };
  uint32_t* const best_s = st.hit_id;
  for (int j = e * 32 + lane; j < QP; j += 128) best_s[j] = 0u;
  epilogue_barrier();
  // ---- pass A: the tile's best upper bound per query ----
#pragma unroll 1
  for (int c = 0; c < QP; c += 16) {
    float u[16];
    load_u(c, u);
    uint32_t mine = 0u;                                                // (see emit_epilogue: one atomic per 16 queries)
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const uint32_t m = __reduce_max_sync(kFull, row_live ? score_code(u[i]) : 0u);
      mine = lane == i ? m : mine;
    }
    if (lane < 16) atomicMax(best_s + c + lane, mine);
  }
  epilogue_barrier();
  for (int j = e * 32 + lane; j < nq_pass; j += 128) p.gmax[size_t(q0 + j) * p.gmax_stride + tile] = best_s[j];
  epilogue_barrier();
  if (e == 0 && lane == 0) {
    grid_arrive(p.gbar + 2 * pass_global);                            // this CTA's bests are published
    grid_wait(p.gbar + 2 * pass_global + 1, uint32_t(n_ctas));        // every CTA's share of the thresholds is
  }
  epilogue_barrier();
  for (int j = e * 32 + lane; j < nq_pass; j += 128) {
    uint32_t code;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(code) : "l"(p.gthr + q0 + j));
    st.thr[j] = code != 0u ? code_score(code) : -INFINITY;            // (dead queries keep +inf)
  }
  epilogue_barrier();
  // ---- pass B: the same accumulators against the thresholds; survivors go straight to the candidate array ----
#pragma unroll 1
  for (int c = 0; c < QP; c += 16) {
    float u[16];
    load_u(c, u);
    uint32_t mask = 0;
#pragma unroll
    for (int i4 = 0; i4 < 4; ++i4) {
      const float4 th4 = *reinterpret_cast<const float4*>(st.thr + c + 4 * i4);
      mask |= (!(u[4 * i4 + 0] < th4.x) ? 1u : 0u) << (4 * i4 + 0);     // NaN passes, like torch.topk ranks it first
      mask |= (!(u[4 * i4 + 1] < th4.y) ? 1u : 0u) << (4 * i4 + 1);
      mask |= (!(u[4 * i4 + 2] < th4.z) ? 1u : 0u) << (4 * i4 + 2);
      mask |= (!(u[4 * i4 + 3] < th4.w) ? 1u : 0u) << (4 * i4 + 3);
    }
    if (!row_live) mask = 0u;
    // (a NaN passes every comparison, also the +inf threshold of a padding query: those have no candidate array)
    if (c + 16 > nq_pass) mask &= c < nq_pass ? (1u << (nq_pass - c)) - 1u : 0u;
    if (mask != 0u) {                                                  // rare
      const uint32_t id = uint32_t(p.idmap(grow));
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if ((mask >> i) & 1u) {
          const int q = q0 + c + i;
          const uint32_t idx = atomicAdd(p.ocount + q, 1u);
          if (idx < uint32_t(p.ocap)) p.ocand[size_t(q) * p.ocap + idx] = pack_key(u[i], id);
        }
      }
    }
  }
  tc_fence_before();
}

__device__ __forceinline__ void oneshot_service(const TcParams& p, const FiltState st, int q0, int nq_pass, int n_ctas, int cta,
                                                int pass_global, float maxn, int lane) {
  if (lane == 0) grid_wait(p.gbar + 2 * pass_global, uint32_t(n_ctas));   // every tile's bests are in gmax
  __syncwarp();
  service_select_round(p, st, q0, nq_pass, n_ctas, cta, maxn, lane, p.k_true);
  __syncwarp();
  if (lane == 0) grid_arrive(p.gbar + 2 * pass_global + 1);
}

// PAIR: two CTAs of one TPC (cluster of 2) run ONE MMA of M = 256 per instruction (cta_group::2): each CTA streams
// its own 128 bank rows and holds HALF of the query operand, which the tensor cores of both read -- 64 instead of
// 96 KB of shared-memory traffic per k-block at 256 queries.  The leader (cluster rank 0) issues the MMAs; both CTAs'
// TMA loads complete on the LEADER's full barrier (cta_group::2 loads, tma_load_2d_pair); commits are multicast to both
// CTAs' barriers; both epilogues release the accumulator on the leader's barrier.  Everything else (thresholds,
// exchange, candidates) is per CTA.
template <int QP, bool MULTI, int MODE, bool PAIR = false>
__global__ void __launch_bounds__(kTcThreads, 1)
scan_tc_kernel(const __grid_constant__ CUtensorMap bank_map, const __grid_constant__ TcParams p) {
  extern __shared__ unsigned char smem_unaligned[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_unaligned) + 1023) & ~uintptr_t(1023));
  constexpr bool BF = MODE == kBf16x3;                   // three bf16 query planes, exact
  constexpr bool BFD = MODE == kBf16x3 || MODE == kBf16x1;   // the bank (and the MMA) is bf16
  constexpr bool FILT = MODE == kTf32x1 || MODE == kBf16x1;
  constexpr bool TWO_ISSUERS = MODE == kTf32x3;
  static_assert(!(FILT && MULTI), "the filter variant accumulates the whole K chain in the tensor core");
  // scan-and-queue epilogue + threshold exchange (filt_epilogue / filt_service): every variant whose tile is one
  // accumulation unit; MULTI (K-sets summed in registers) keeps the score-tile epilogue below
  constexpr bool SCANEPI = !MULTI;
  constexpr int kParts = BF ? 3 : (FILT ? 1 : 2);        // row groups of the query operand per k-block
  constexpr int kElemsK = BFD ? 64 : 32;                 // bank columns per k-block (always 128 bytes)
  static_assert(!PAIR || (FILT && QP >= 64), "CTA pairs: filter variant, 64+ queries per pass");
  // (emit and one-shot mode keep no per-CTA lists: the space goes to ring stages, see scan_tc_plan)
  const int list_cap = (FILT && (p.emit || p.oneshot)) ? 0 : p.kcap;
  const bool bres = FILT && p.b_resident != 0;
  const SmemLayout L = tc_layout(p.stages, p.qp, list_cap, kParts, PAIR, bres ? p.n_kblocks : 0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.stages;
  constexpr int kBRows = PAIR ? QP / 2 : QP;             // query rows of the operand in THIS CTA's shared memory
  const uint32_t b_stage_bytes = uint32_t(kParts * kBRows) * 128;
  const uint32_t pair_rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = pair_rank == 0u;

  uint64_t* lists = reinterpret_cast<uint64_t*>(smem + L.lists);
  float* norms = reinterpret_cast<float*>(smem + L.norms);
  uint64_t* full_tma = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* full_split = full_tma + kMaxStages;
  uint64_t* empty = full_split + kMaxStages;
  uint64_t* peer_full = empty + kMaxStages;        // [0]: the resident query operand has landed (b_resident); others unused
  uint64_t* tmem_full = peer_full + kMaxStages;    // [kMaxAccBufs]
  uint64_t* tmem_empty = tmem_full + kMaxAccBufs;  // [kMaxAccBufs]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.tmem_slot);

  // this CTA's share of the work
  const int pass_local = int(blockIdx.x) % p.passes;
  const int tile0 = int(blockIdx.x) / p.passes, tile_step = int(gridDim.x) / p.passes;
  const int q0 = p.q0 + pass_local * QP;
  const int nq_pass = min(QP, p.n_queries - q0);
  pdl_wait();                                            // (programmatic dependent launch: the previous kernel's results)
  if (p.need) {                                          // conditional (fallback) launch: nothing to do for most passes
    int any = 0;
    for (int i = threadIdx.x; i < nq_pass; i += kTcThreads) any |= int(p.need[q0 + i] != 0u);
    if (!__syncthreads_or(any)) return;
  }

  // ---- one-time setup ------------------------------------------------------------------------
  for (int i = threadIdx.x; i < QP * list_cap; i += kTcThreads) lists[i] = kEmptyKey;
  FiltState st{};
  float maxn = 1.f;                                      // un-normalised scores: E scales with ||row|| <= maxn
  if constexpr (SCANEPI) {
    float* fs = reinterpret_cast<float*>(smem + L.thr);
    st.eq = fs; st.thr = fs + QP; st.lost = fs + 2 * QP; st.thr_ext = fs + 3 * QP; st.qs = fs + 4 * QP;
    st.qn = reinterpret_cast<uint32_t*>(fs + 5 * QP);
    st.hit_u = reinterpret_cast<float*>(smem + L.scratch);
    st.hit_id = reinterpret_cast<uint32_t*>(smem + L.scratch) + (list_cap ? hit_cap(QP) : 0);   // (list-free: best / published / count words only)
    if (FILT && !(p.flags & DEVQA_NORMALIZE_BANK)) maxn = sqrtf(__uint_as_float(*p.bank_max_sumsq));
    for (int j = threadIdx.x; j < QP; j += kTcThreads) {
      const bool live = j < nq_pass;
      const float eq = (FILT && live) ? p.qerr[q0 + j] : 0.f;
      float t = live ? -INFINITY : INFINITY;                          // dead queries never hit
      const uint32_t code = (live && p.seed_u) ? p.seed_u[q0 + j] : 0u;
      if (code != 0u) {
        const float ts = code_score(code) - 2.f * eq * maxn;          // lower bound of the exact k-th best score
        if (ts > t) t = ts;
      }
      st.eq[j] = eq; st.thr[j] = t; st.lost[j] = -INFINITY; st.thr_ext[j] = -INFINITY;
      st.qs[j] = (FILT && live && p.qscale) ? p.qscale[q0 + j] : 1.f;
    }
    if (threadIdx.x < 4) st.qn[threadIdx.x] = 0u;
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&bank_map);
  }
  // bf16 banks with cached row norms need no pass over the tile by CUDA cores at all; the filter variant
  // needs ||row||^2 either way (normalised scores, or max ||row|| for the error bound of plain dot scores)
  const bool norm_warps = FILT ? p.row_sumsq == nullptr
                               : !(BF && (p.row_sumsq != nullptr || !(p.flags & DEVQA_NORMALIZE_BANK)));
  // emit mode, 128+ queries per pass, cached norms: warps 4-7 are a second group of epilogue warps (emit_epilogue)
  const int epi_groups = (FILT && QP >= 128 && p.emit && !norm_warps && !(p.flags & kDbgOneEpiGroup)) ? 2 : 1;
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_tma[s], 1); mbar_init(&full_split[s], 128); mbar_init(&empty[s], TWO_ISSUERS ? 2 : 1);
      mbar_init(&peer_full[s], 1);
    }
    for (int a = 0; a < kMaxAccBufs; ++a) {
      mbar_init(&tmem_full[a], TWO_ISSUERS ? 2 : 1);
      mbar_init(&tmem_empty[a], uint32_t((PAIR ? 256 : 128) * epi_groups));
    }
    fence_barrier_init();
  }
  if (warp == 2) { if constexpr (PAIR) tmem_alloc_pair(tmem_slot, 512); else tmem_alloc(tmem_slot, 512); }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();   // (pairs: the peer's barriers are initialised too)
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const float* const qcat = p.qcat + size_t(pass_local) * size_t(p.n_kblocks) * (kParts * QP * 32);   // 32 words per 128-byte row
  // TMEM columns: two accumulator buffers of 3*QP (filter: QP) columns, then one 32-column lo tile per ring stage
  constexpr uint32_t kAccCols = FILT ? QP : 3 * QP;
  const uint32_t tmem_alo = tmem_base + 2 * kAccCols;
  const int nkb = p.n_kblocks;
  // Tile sequence of this CTA: its strided tiles, then (filter variant, 8+ tiles) the first tile AGAIN: on its
  // first visit the epilogue only learns each query's best upper bound from it (thresholds for everyone through
  // the service warps) and waits a bounded moment for them; on the second visit it is a tile like any other --
  // with thresholds in place.  Without this, the first two tiles of every CTA cost as much as ~40 ordinary ones
  // (lists built from nothing, one insertion at a time).
  // (pairs: both CTAs run the leader's count; a missing last tile of the peer is all out-of-bounds rows)
  const int n_my = (p.n_tiles - (PAIR ? (tile0 & ~1) : tile0) + tile_step - 1) / tile_step;
  const int boot = (SCANEPI && n_my >= 8 && tile_step >= p.k_true && p.gmax && !(p.flags & kDbgNoBoot) && !(FILT && p.emit)) ? 1 : 0;
  const int n_iter = n_my + boot;
  auto tile_at = [&](int it) { return tile0 + (it < n_my ? it : it - n_my) * tile_step; };
  // Accumulator buffers.  Two everywhere -- except in emit mode, where the epilogue's first tile waits ~20 us for the
  // first threshold exchange: with 512 / QP buffers (8 at 64 queries per pass) the MMA pipeline runs up to seven tiles
  // ahead instead of stalling behind that wait, and the epilogue (2.7 k cycles per tile against a 7 k tile period)
  // catches up afterwards.
  // (in-kernel norm warps keep a ring of kNormRing tiles of row norms: the epilogue must not fall further behind)
  const int n_acc = (FILT && p.emit && !(p.flags & kDbgNoBoot)) ? min(norm_warps ? 4 : kMaxAccBufs, 512 / QP) : 2;
  const int set_kb = FILT ? nkb : kSetKBlocks;           // k-blocks accumulated inside the tensor core
  const int nsets = (nkb + set_kb - 1) / set_kb;         // accumulation units per tile (1 unless MULTI)

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      // (An L2 prefetch of the tiles a dozen k-blocks ahead of the ring used to run here.  With 4-8 ring stages in
      // flight it only cost: every byte touched L2 twice.  Measured at 10M x 768, sustained: Q = 16 5.29 -> 4.40 ms,
      // Q = 64 5.19 -> 4.90 ms, Q = 1024 31.8 -> 28.5 ms, 1M rows Q = 64 0.594 -> 0.524 ms.  Opt-in for experiments.)
      const bool pf_on = (p.flags & kDbgL2Prefetch) != 0;
      const bool qtma = !(p.flags & kDbgNoQueryTma);
      const uint32_t tx_bytes = kATileBytes + ((qtma && !bres) ? b_stage_bytes : 0u);
      if (bres && qtma) {
        // the query operand, once: every k-block of the pass (at 64 queries x 768 fp16 columns: 96 KB).  Streaming it
        // with the stages made each SM pull 1.5x the bank's bytes out of L2 -- under the power cap (SM clock ~1.45 GHz)
        // that, not HBM, set the tile period.
        if constexpr (PAIR) {
          // each CTA its half of the rows; both halves are counted on the LEADER's barrier (it issues the MMAs)
          const uint32_t rb = mapa_u32(smem_u32(&peer_full[0]), 0u);
          if (leader) mbar_expect_tx(&peer_full[0], 2u * uint32_t(nkb) * b_stage_bytes);
          for (int kb = 0; kb < nkb; ++kb)
            tma_load_2d_pair(smem + L.bq + kb * b_stage_bytes, &p.qmap, rb, 0,
                             (pass_local * nkb + kb) * (kParts * QP) + int(pair_rank) * kBRows);
        } else {
          mbar_expect_tx(&peer_full[0], uint32_t(nkb) * b_stage_bytes);
          for (int kb = 0; kb < nkb; ++kb)
            bulk_load(smem + L.bq + kb * b_stage_bytes, qcat + size_t(kb) * (kParts * QP * 32), b_stage_bytes, &peer_full[0]);
        }
      }
      // (a bank that about fits in L2 should stay there between calls and between the passes of a call: C2 is 67-134 MB)
      const bool contig = (p.flags & kDbgContigTiles) != 0;            // (timing experiment: see scan_tc_launch)
      const bool ef = !(p.flags & kDbgNoEvictFirst) && !p.oneshot && int64_t(p.n_tiles) * nkb * int64_t(kATileBytes) > (int64_t(192) << 20);
      const uint64_t pol = l2_policy_evict_first();
      const uint32_t leader_full0 = PAIR ? mapa_u32(smem_u32(&full_tma[0]), 0u) : 0u; (void)leader_full0;
      int pf_it = 0, pf_kb = 0;                           // next k-block to prefetch
      auto prefetch_next = [&]() {
        if (pf_it < n_iter) {
          if (pf_on) tma_prefetch_l2_2d(&bank_map, pf_kb * kElemsK, tile_at(pf_it) * kTileRows);
          if (++pf_kb == nkb) { pf_kb = 0; ++pf_it; }
        }
      };
      for (int i = 0; i < kPrefetchAhead; ++i) prefetch_next();
      RingPos r;
      uint32_t trace_it = 0;
      for (int it = 0; it < n_iter; ++it) {
        const int tile = tile_at(it); (void)tile;
        for (int kb = 0; kb < nkb; ++kb) {
          prefetch_next();
          mbar_wait(&empty[r.s], r.phase ^ 1u);
          trace_at(p, 1, trace_it);
          if constexpr (PAIR) {
            // both CTAs' bytes complete on the leader's barrier (the leader expects twice a stage)
            const uint32_t lbar = leader_full0 + uint32_t(r.s) * 8u;
            if (leader) mbar_expect_tx(&full_tma[r.s], 2u * tx_bytes);
            if (ef) tma_load_2d_pair_hint(smem + L.a_raw + r.s * kATileBytes, &bank_map, lbar, contig ? 0 : kb * kElemsK, contig ? (tile * nkb + kb) * kTileRows : tile * kTileRows, pol);
            else tma_load_2d_pair(smem + L.a_raw + r.s * kATileBytes, &bank_map, lbar, contig ? 0 : kb * kElemsK, contig ? (tile * nkb + kb) * kTileRows : tile * kTileRows);
            if (qtma && !bres)   // this CTA's half of the k-block's query rows (pre-swizzled rows of 128 bytes)
              tma_load_2d_pair(smem + L.bq + r.s * b_stage_bytes, &p.qmap, lbar, 0,
                               (pass_local * nkb + kb) * (kParts * QP) + int(pair_rank) * kBRows);
          } else {
            mbar_expect_tx(&full_tma[r.s], tx_bytes);
            if (ef) tma_load_2d_hint(smem + L.a_raw + r.s * kATileBytes, &bank_map, &full_tma[r.s], contig ? 0 : kb * kElemsK, contig ? (tile * nkb + kb) * kTileRows : tile * kTileRows, pol);
            else tma_load_2d(smem + L.a_raw + r.s * kATileBytes, &bank_map, &full_tma[r.s], contig ? 0 : kb * kElemsK, contig ? (tile * nkb + kb) * kTileRows : tile * kTileRows);
            if (qtma && !bres)     // the whole [2QP x 32] query k-block is contiguous (and pre-swizzled) in global memory
              bulk_load(smem + L.bq + r.s * b_stage_bytes, qcat + size_t(kb) * (kParts * QP * 32), b_stage_bytes, &full_tma[r.s]);
          }
          trace_at(p, 0, trace_it++);
          r.advance(S);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer 1: A (hi by truncation, from shared memory) x [Qhi ; Qlo] -> cols [0, 2QP) =====
    if (PAIR && !leader) {
      // pair, peer CTA: no MMA issue (its loads complete on the leader's barrier)
    } else if (lane == 0) {
      const bool f16 = BFD && (p.flags & kFlagShadowScan) != 0;       // the shadow of an fp32 bank is (scaled) fp16, not bf16
      const uint32_t idesc_cat = BFD ? (f16 ? (PAIR ? idesc_f16_pair(QP) : idesc_f16(kParts * QP))
                                            : (PAIR ? idesc_bf16_pair(QP) : idesc_bf16(kParts * QP)))
                                     : (PAIR ? idesc_tf32_pair(QP) : idesc_tf32(kParts * QP));
      const uint32_t a_lo0 = (smem_u32(smem + L.a_raw) & 0x3ffffu) >> 4, b_lo0 = (smem_u32(smem + L.bq) & 0x3ffffu) >> 4;
      RingPos r;
      uint32_t unit = 0;
      if (bres && !(p.flags & kDbgNoQueryTma)) mbar_wait(&peer_full[0], 0);   // the resident query operand has landed
      for (int it = 0; it < n_iter; ++it) {
        const int tile = tile_at(it); (void)tile;
        for (int set = 0; set < nsets; ++set, ++unit) {
          const int a = int(unit % uint32_t(n_acc));
          const uint32_t a_par = (unit / uint32_t(n_acc)) & 1u;
          if constexpr (PAIR) mbar_wait_cluster(&tmem_empty[a], a_par ^ 1);   // both CTAs' epilogues
          else mbar_wait(&tmem_empty[a], a_par ^ 1);
          tc_fence_after();
          const uint32_t d = tmem_base + uint32_t(a) * kAccCols;
          const int kb_end = min(nkb, (set + 1) * set_kb);
          for (int kb = set * set_kb; kb < kb_end; ++kb) {
            const uint32_t first = kb == set * set_kb ? 0u : 1u;
            const uint32_t a_lo = a_lo0 + uint32_t(r.s) * (kATileBytes >> 4), b_lo = b_lo0 + uint32_t(bres ? kb : r.s) * (b_stage_bytes >> 4);
            mbar_wait(&full_tma[r.s], r.phase);
            trace_at(p, 3, unit * uint32_t(nkb) + uint32_t(kb));
            trace_at(p, 4, unit * uint32_t(nkb) + uint32_t(kb));            // (pairs: the barrier counts both CTAs' bytes)
            tc_fence_after();
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {              // four 32-byte K slices of the 128-byte rows
              if (p.flags & kDbgNoMma1) break;            // (bisection: stream the tiles only)
              if constexpr (BFD && PAIR) mma_bf16_pair(d, desc_from_lo(a_lo + kk * 2), desc_from_lo(b_lo + kk * 2), idesc_cat, first | uint32_t(kk));
              else if constexpr (BFD) mma_bf16(d, desc_from_lo(a_lo + kk * 2), desc_from_lo(b_lo + kk * 2), idesc_cat, first | uint32_t(kk));
              else if constexpr (PAIR) mma_tf32_pair(d, desc_from_lo(a_lo + kk * 2), desc_from_lo(b_lo + kk * 2), idesc_cat, first | uint32_t(kk));
              else mma_tf32(d, desc_from_lo(a_lo + kk * 2), desc_from_lo(b_lo + kk * 2), idesc_cat, first | uint32_t(kk));
            }
            if (!TWO_ISSUERS && norm_warps) mbar_wait(&full_split[r.s], r.phase);   // the norm warps are done reading the stage
            if constexpr (PAIR) mma_commit_pair(&empty[r.s]);      // frees the stage in both CTAs
            else mma_commit(&empty[r.s]);                 // (one of) the arrival(s) that free the stage
            r.advance(S);
          }
          if constexpr (PAIR) mma_commit_pair(&tmem_full[a]); else mma_commit(&tmem_full[a]);
        }
      }
    }
  } else if (warp == 2) {
    // ===== MMA issuer 2: Alo (from tensor memory) x Qhi -> cols [2QP, 3QP) =====
    if (lane == 0 && TWO_ISSUERS) {
      const uint32_t idesc_hi = idesc_tf32(QP);
      const uint32_t b_lo0 = (smem_u32(smem + L.bq) & 0x3ffffu) >> 4;
      const bool on = !(p.flags & kDbgNoMma2);
      RingPos r;
      uint32_t unit = 0;
      for (int it = 0; it < n_iter; ++it) {
        const int tile = tile_at(it); (void)tile;
        for (int set = 0; set < nsets; ++set, ++unit) {
          const int a = unit & 1;
          mbar_wait(&tmem_empty[a], ((unit >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d = tmem_base + uint32_t(a) * kAccCols + 2 * QP;
          const int kb_end = min(nkb, (set + 1) * set_kb);
          for (int kb = set * set_kb; kb < kb_end; ++kb) {
            const uint32_t first = kb == set * set_kb ? 0u : 1u;
            const uint32_t b_lo = b_lo0 + uint32_t(r.s) * (b_stage_bytes >> 4);
            const uint32_t alo = tmem_alo + uint32_t(r.s) * kBlockK;
            mbar_wait(&full_split[r.s], r.phase);         // implies full_tma[s]: the split warps waited on it
            tc_fence_after();
            if (on) {
#pragma unroll
              for (int kk = 0; kk < kBlockK / 8; ++kk)
                mma_tf32_ts(d, alo + kk * 8, desc_from_lo(b_lo + kk * 2), idesc_hi, first | uint32_t(kk));
            }
            mma_commit(&empty[r.s]);
            r.advance(S);
          }
          mma_commit(&tmem_full[a]);
        }
      }
    }
  } else if (warp >= 4 && warp < 8 && norm_warps) {
    // ===== split warps: thread <-> bank row of the tile =====
    const int row = threadIdx.x - 128;
    const uint32_t row_off = uint32_t(row) * 128;
    const uint32_t sw = uint32_t(row & 7);                // 16-byte chunk c of row r lives at chunk c ^ (r & 7)
    uint32_t lt = 0;
    RingPos rp;
    const uint32_t st_lane = uint32_t((warp & 3) * 32) << 16;     // warp w may only touch TMEM lanes 32*(w%4)..
    const uint32_t a_raw_s = smem_u32(smem + L.a_raw) + row_off;
    for (int it = 0; it < n_iter; ++it, ++lt) {
      const int tile = tile_at(it); (void)tile;
      float ss[8];                                        // one partial sum per 16-byte chunk position
#pragma unroll
      for (int c = 0; c < 8; ++c) ss[c] = 0.f;
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = rp.s;
        mbar_wait(&full_tma[s], rp.phase);
        const uint32_t src = a_raw_s + uint32_t(s) * kATileBytes;     // shared-space address of this thread's row
        if constexpr (BF) {
          if (p.flags & DEVQA_NORMALIZE_BANK) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const float4 v = lds128(src + (uint32_t(c) << 4));      // order is irrelevant for a sum of squares
              const uint32_t w[4] = {__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w)};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float e0 = __uint_as_float(w[i] << 16), e1 = __uint_as_float(w[i] & 0xffff0000u);
                ss[c] = fmaf(e0, e0, ss[c]);
                ss[c] = fmaf(e1, e1, ss[c]);
              }
            }
          }
        }
        if constexpr (FILT && !BFD) {                     // (the bf16 filter runs on banks with cached norms only)
#pragma unroll
          for (int c = 0; c < 8; ++c) {                   // order is irrelevant for a sum of squares
            const float4 v = lds128(src + (uint32_t(c) << 4));
            ss[c] = fmaf(v.x, v.x, ss[c]); ss[c] = fmaf(v.y, v.y, ss[c]);
            ss[c] = fmaf(v.z, v.z, ss[c]); ss[c] = fmaf(v.w, v.w, ss[c]);
          }
        }
        uint32_t lo[32];
        if (TWO_ISSUERS && !(p.flags & kDbgNoSplit)) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {                   // logical 16-byte chunk c of the row
            const float4 v = lds128(src + ((uint32_t(c) ^ sw) << 4));
            lo[4 * c + 0] = __float_as_uint(v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u));
            lo[4 * c + 1] = __float_as_uint(v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u));
            lo[4 * c + 2] = __float_as_uint(v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u));
            lo[4 * c + 3] = __float_as_uint(v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u));
            ss[c] = fmaf(v.x, v.x, ss[c]); ss[c] = fmaf(v.y, v.y, ss[c]);
            ss[c] = fmaf(v.z, v.z, ss[c]); ss[c] = fmaf(v.w, v.w, ss[c]);
          }
          // lane (= row) x 32 columns of this stage's lo tile; warp w may only touch TMEM lanes 32*(w%4)..
          tmem_st32(tmem_alo + uint32_t(s) * kBlockK + st_lane, lo);
          tmem_st_wait();
        }
        if (kb == nkb - 1) {
          const float tot = ((ss[0] + ss[1]) + (ss[2] + ss[3])) + ((ss[4] + ss[5]) + (ss[6] + ss[7]));
          norms[(lt % kNormRing) * kTileRows + row] = tot;
        }
        tc_fence_before();                                // TMEM writes ordered before the MMA issuer's reads
        mbar_arrive(&full_split[s]);
        rp.advance(S);
      }
    }
  } else if (SCANEPI && warp == 3) {
    // ===== service warp: threshold exchange between CTAs =====
    if (FILT && p.oneshot) oneshot_service(p, st, q0, nq_pass, tile_step, tile0, p.q0 / QP + pass_local, maxn, lane);
    else filt_service(p, st, q0, nq_pass, tile_step, tile0, maxn, lane);
  } else if (warp >= 8 && FILT && p.oneshot) {
    // ===== epilogue warps, one tile per CTA: two passes over the accumulators, candidates straight to global memory =====
    if constexpr (FILT)
      oneshot_epilogue<QP>(p, st, norms, norm_warps, tmem_base, tmem_full, tile0, tile_step, p.q0 / QP + pass_local, q0, nq_pass,
                           warp - 8, lane);
  } else if (FILT && p.emit && (warp >= 8 || (warp >= 4 && epi_groups == 2))) {
    // ===== epilogue warps, long scan without lists: survivors straight to the global candidate array =====
    if constexpr (FILT)
      emit_epilogue<QP>(p, st, norms, norm_warps, tmem_base, tmem_full, tmem_empty, tile0, tile_step, n_my, q0, nq_pass, warp & 3, lane,
                        n_acc, (PAIR && !leader) ? mapa_u32(smem_u32(&tmem_empty[0]), 0u) : 0u, warp >= 8 ? 0 : 1, epi_groups);
  } else if (warp >= 8 && SCANEPI) {
    // ===== epilogue warps: scan + hit queue (filt_epilogue) =====
    filt_epilogue<QP, FILT ? 1 : 3>(p, st, lists, norms, norm_warps, tmem_base, tmem_full, tmem_empty, tile0, tile_step, n_my, boot,
                                    q0, nq_pass, maxn, warp - 8, lane,
                                    (PAIR && !leader) ? mapa_u32(smem_u32(&tmem_empty[0]), 0u) : 0u);
  } else if (warp >= 8) {
    // ===== epilogue warps =====
    // Phase A: thread <-> TMEM lane <-> bank row: sum the three accumulator parts, scale by 1/||row||
    //          and drop the scores into a [query][row] shared-memory tile (HQ queries at a time).
    // Phase B: warp e exclusively owns queries e*HQ/4 .. of the tile: it scans their 128 scores (one
    //          float4 per lane), filters them against the running k-th score and inserts survivors into
    //          the query's sorted list.  No locks; two 128-thread named barriers per half tile.
    constexpr int HQ = QP < 32 ? QP : 32;                 // queries per score-tile fill
    constexpr int HALVES = QP / HQ;
    const int e = warp - 8;                               // TMEM lane quarter = warp % 4
    const int row_in_tile = e * 32 + lane;
    float* stile = reinterpret_cast<float*>(smem + L.scratch);        // [HQ][128]
    const bool no_insert = (p.flags & kDbgNoInsert) != 0;
    int slow_calls = 0;
    float thr_reg = -INFINITY;                            // lane l: running k-th score of the l-th query this warp owns
    uint32_t lt = 0, unit = 0;
    // lane l of warp e owns query  own_j(l) = (l / NQW_) * HQ + e * NQW_ + l % NQW_   (l < QP / 4)
    constexpr int NQW_ = HQ / 4;
    const int own_j = (lane / NQW_) * HQ + e * NQW_ + lane % NQW_;
    const bool own_live = lane < QP / 4 && own_j < nq_pass;
    uint32_t* const own_gthr = p.gthr + q0 + (own_live ? own_j : 0);
    for (int it = 0; it < n_iter; ++it, ++lt) {
      const int tile = tile_at(it); (void)tile;
      // pick up the bounds other CTAs have published (one relaxed load per owned query per tile)
      if (own_live) {
        uint32_t code;
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(code) : "l"(own_gthr));
        if (code != 0u) thr_reg = fmaxf(thr_reg, code_score(code));
      }
      float acc[MULTI ? QP : 1];                          // fp32 RN sum over the tile's K-sets
      for (int set = 0; set < nsets; ++set, ++unit) {
        const int a = unit & 1;
        const bool last = set == nsets - 1;
        mbar_wait(&tmem_full[a], (unit >> 1) & 1);
        tc_fence_after();
        if (e == 0 && lane == 0) trace_at(p, 6, unit);
        float inv = 1.f;
        if (last && (p.flags & DEVQA_NORMALIZE_BANK)) {
          const int64_t grow = int64_t(tile) * kTileRows + row_in_tile;
          const float ssq = norm_warps ? norms[(lt % kNormRing) * kTileRows + row_in_tile]
                                       : (grow < p.n_rows ? p.row_sumsq[grow] : 1.f);
          const float nrm = sqrtf(ssq);
          inv = 1.0f / ((p.flags & DEVQA_NOEPS) ? nrm : fmaxf(nrm, kNormEps));
        }
        const uint32_t taddr = tmem_base + uint32_t(a) * kAccCols + (uint32_t(e * 32) << 16);
#pragma unroll (MULTI ? HALVES : 1)
        for (int h = 0; h < HALVES; ++h) {
          // ---- phase A ----
#pragma unroll
          for (int c = 0; c < HQ / 16; ++c) {
            const int j0 = h * HQ + c * 16;
            uint32_t d1[16], d2[16], d3[16];
            tmem_ld<16>(taddr + uint32_t(j0), d1);
            tmem_ld<16>(taddr + uint32_t(QP + j0), d2);
            tmem_ld<16>(taddr + uint32_t(2 * QP + j0), d3);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float v = (__uint_as_float(d2[i]) + __uint_as_float(d3[i])) + __uint_as_float(d1[i]);
              if constexpr (MULTI) {
                v = set == 0 ? v : acc[j0 + i] + v;
                acc[j0 + i] = v;
              }
              if (last) stile[(c * 16 + i) * kTileRows + row_in_tile] = v * inv;
            }
          }
          if (h == HALVES - 1) {                          // every TMEM read of this buffer has completed
            tc_fence_before();
            mbar_arrive(&tmem_empty[a]);
          }
          if (!last) continue;
          epilogue_barrier();
          // ---- phase B ----
          const int64_t row0 = int64_t(tile) * kTileRows + lane * 4;
          uint32_t rowmask = 0;
#pragma unroll
          for (int t = 0; t < 4; ++t) rowmask |= (row0 + t < p.n_rows ? 1u : 0u) << t;
          if (no_insert) rowmask = 0;
          constexpr int NQW = HQ / 4;                     // queries this warp owns per half
          float4 sc[NQW];
#pragma unroll
          for (int i = 0; i < NQW; ++i)                   // all tile reads in flight together
            sc[i] = *reinterpret_cast<const float4*>(stile + (e * NQW + i) * kTileRows + lane * 4);
          uint32_t anyq = 0;
#pragma unroll
          for (int i = 0; i < NQW; ++i) {
            const float th = __shfl_sync(kFull, thr_reg, h * NQW + i);
            const bool hit = (!(sc[i].x < th) && (rowmask & 1u)) || (!(sc[i].y < th) && (rowmask & 2u)) ||
                             (!(sc[i].z < th) && (rowmask & 4u)) || (!(sc[i].w < th) && (rowmask & 8u));
            anyq |= (__ballot_sync(kFull, hit) ? 1u : 0u) << i;
          }
          while (anyq) {                                  // warp-uniform
            ++slow_calls;
            const int i = __ffs(anyq) - 1;
            anyq &= anyq - 1;
            const int j = h * HQ + e * NQW + i;
            if (j >= nq_pass) break;
            const int oi = h * NQW + i;                   // index among the queries this warp owns
            const float th = __shfl_sync(kFull, thr_reg, oi);
            const float nt = epilogue_insert_query(lists + size_t(j) * p.kcap, p.kcap, p.top_k,
                                                   stile + (e * NQW + i) * kTileRows, th, rowmask, row0, p.idmap, lane);
            if (lane == oi && nt > thr_reg) {               // publish: a lower bound of the global k-th score
              thr_reg = nt;
              atomicMax(own_gthr, score_code(nt));
            }
          }
          epilogue_barrier();                             // the score tile may be overwritten again
        }
        if (e == 0 && lane == 0) trace_at(p, 7, unit);
#ifdef DEVQA_TC_TRACE
        if (e == 0 && lane == 0 && p.trace && blockIdx.x == 0 && unit < 256) { p.trace[8 * 256 + unit] = slow_calls; slow_calls = 0; }
#endif
      }
    }
  }

  // ---- teardown: candidates out, TMEM back ------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  for (int q = warp; q < nq_pass && !p.oneshot && !p.emit; q += kTcThreads / 32) {
    uint64_t* out = p.cand + (size_t(tile0) * p.n_queries + (q0 + q)) * p.kcap;
    const uint64_t* src = lists + size_t(q) * p.kcap;
    for (int i = lane; i < p.kcap; i += 32) out[i] = src[i];
  }
  if constexpr (PAIR) {
    cluster_sync_all();                                  // nobody leaves (or frees TMEM) while the other still computes
    if (warp == 2) tmem_dealloc_pair(tmem_base, 512);
  } else {
    if (warp == 2) tmem_dealloc(tmem_base, 512);
  }
}

// ---- query preparation: normalise, pad, split into TF32 hi / lo, lay out as the MMA B operand -------------
// Output per pass of qp queries: [n_kblocks][2*qp rows][32 cols] fp32, rows 0..qp-1 = hi, qp..2qp-1 = lo,
// each row's eight 16-byte chunks permuted by the 128-byte swizzle (chunk ^= row & 7), so that a k-block
// is one contiguous, shared-memory-ready copy.  One warp per query.
__global__ void __launch_bounds__(256) prep_queries_kernel(const float* __restrict__ q, int64_t q_ld, int n_queries, int dim,
                                                           int qp, int n_kblocks, int q_pad, unsigned flags,
                                                           float* __restrict__ qcat, uint32_t* __restrict__ gthr,
                                                           const uint32_t* __restrict__ need, uint32_t* __restrict__ gmax,
                                                           int gmax_stride) {
  pdl_wait();
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < q_pad; i += blockDim.x) gthr[i] = 0u;
  if (gmax)                                              // threshold exchange: no CTA has published anything yet
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < q_pad * gmax_stride; i += gridDim.x * blockDim.x) gmax[i] = 0u;
  const int lane = threadIdx.x & 31;
  const int warps = gridDim.x * (blockDim.x >> 5);
  const int d_pad = n_kblocks * kBlockK;
  for (int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < q_pad; r += warps) {
    if (need) {                                          // conditional (fallback) run: only passes with a flagged query
      const int p0 = r / qp * qp;
      bool any = false;
      for (int i = p0 + lane; i < p0 + qp && i < n_queries; i += 32) any |= need[i] != 0u;
      if (!__any_sync(kFull, any)) continue;
    }
    float inv = 0.f;
    const float* src = q + size_t(r < n_queries ? r : 0) * q_ld;
    if (r < n_queries) {
      inv = 1.f;
      if (flags & DEVQA_NORMALIZE_QUERY) {
        float ss = 0.f;
        for (int c = lane; c < dim; c += 32) { const float x = src[c]; ss = fmaf(x, x, ss); }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) ss += __shfl_xor_sync(kFull, ss, off);
        const float nrm = sqrtf(ss);
        inv = 1.0f / ((flags & DEVQA_NOEPS) ? nrm : fmaxf(nrm, kNormEps));
      }
    }
    const int pass = r / qp, rp = r % qp;                 // row rp of the pass: hi at rp, lo at qp + rp
    float* base = qcat + size_t(pass) * n_kblocks * (2 * qp * kBlockK);
    for (int c = lane; c < d_pad; c += 32) {
      const float x = (r < n_queries && c < dim) ? src[c] * inv : 0.f;
      uint32_t hb, lb;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(x));
      const float hi = __uint_as_float(hb);
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(x - hi));
      const int kb = c / kBlockK, w = c % kBlockK;
      float* blk = base + size_t(kb) * (2 * qp * kBlockK);
      const int hi_row = rp, lo_row = qp + rp;
      blk[hi_row * kBlockK + (((w >> 2) ^ (hi_row & 7)) << 2) + (w & 3)] = hi;
      blk[lo_row * kBlockK + (((w >> 2) ^ (lo_row & 7)) << 2) + (w & 3)] = __uint_as_float(lb);
    }
  }
}

// Filter variants (kTf32x1 / kBf16x1): ONE plane, rounded to nearest -- TF32 words [n_kblocks][qp rows][32 cols] or bf16
// [n_kblocks][qp rows][64 cols] per pass -- plus the per-query scale and norm that refine.cu needs to re-create the exact
// fp32 query and the error bound.  One CTA per (padded) query: the row is staged in shared memory by all threads
// (full memory-level parallelism; one warp walking a 4096-column query was 13.5 us of C2's 240), warp 0 then forms the
// norms from shared memory in EXACTLY the stream kernel's order (lane-strided FMAs, xor butterfly: the scale has to be
// bit-identical to scan_stream.cu's for the re-scored results to be), and all threads write the operand.
// Also zeroes, grid-wide: gthr / seed_u [q_pad], the threshold exchange table gmax, and `zero_words` (one-shot
// mode: arrival counters + candidate counts).
template <bool BF16>
__global__ void __launch_bounds__(128) prep_queries_filt_kernel(const float* __restrict__ q, int64_t q_ld, int n_queries, int dim,
                                                                int qp, int n_kblocks, int q_pad, unsigned flags,
                                                                void* __restrict__ qcat_, uint32_t* __restrict__ gthr,
                                                                float* __restrict__ qinv, float* __restrict__ qerr,
                                                                uint32_t* __restrict__ seed_u, uint32_t* __restrict__ gmax,
                                                                int gmax_stride, uint32_t* __restrict__ zero_words, int n_zero,
                                                                float* __restrict__ qscale) {
  extern __shared__ __align__(16) float srow[];          // [d_pad]
  __shared__ float s_inv, s_up1, s_up2;
  constexpr int EK = BF16 ? 64 : 32;                     // columns per 128-byte k-block row
  const int tid = threadIdx.x, nthr = blockDim.x;
  pdl_wait();
  for (int i = blockIdx.x * nthr + tid; i < q_pad; i += gridDim.x * nthr) { gthr[i] = 0u; seed_u[i] = 0u; }
  if (gmax)                                              // threshold exchange: no CTA has published anything yet
    for (int i = blockIdx.x * nthr + tid; i < q_pad * gmax_stride; i += gridDim.x * nthr) gmax[i] = 0u;
  for (int i = blockIdx.x * nthr + tid; i < n_zero; i += gridDim.x * nthr) zero_words[i] = 0u;
  const int lane = tid & 31, warp = tid >> 5;
  const int d_pad = n_kblocks * EK;
  for (int r = blockIdx.x; r < q_pad; r += gridDim.x) {
    const bool live = r < n_queries;
    const float* src = q + size_t(live ? r : 0) * q_ld;
    __syncthreads();
    if (live && (dim & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {   // 128-bit loads, 8 in flight per thread
      const float4* src4 = reinterpret_cast<const float4*>(src);
      float4* dst4 = reinterpret_cast<float4*>(srow);
#pragma unroll 8
      for (int c = tid; c < dim / 4; c += nthr) dst4[c] = __ldg(src4 + c);
      for (int c = dim + tid; c < d_pad; c += nthr) srow[c] = 0.f;
    } else {
#pragma unroll 8
      for (int c = tid; c < d_pad; c += nthr) srow[c] = (live && c < dim) ? src[c] : 0.f;
    }
    __syncthreads();
    if (warp == 0) {
      float inv = 0.f;
      if (live) {
        inv = 1.f;
        if (flags & DEVQA_NORMALIZE_QUERY) {
          float ss = 0.f;
          for (int c = lane; c < dim; c += 32) { const float x = srow[c]; ss = fmaf(x, x, ss); }
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1) ss += __shfl_xor_sync(kFull, ss, off);
          const float nrm = sqrtf(ss);
          inv = 1.0f / ((flags & DEVQA_NOEPS) ? nrm : fmaxf(nrm, kNormEps));
        }
      }
      float nn = 0.f, mx = 0.f;                          // ||query as used||^2, its largest magnitude
      for (int c = lane; c < d_pad; c += 32) { const float x = srow[c] * inv; nn = fmaf(x, x, nn); mx = fmaxf(mx, fabsf(x)); }
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) {
        nn += __shfl_xor_sync(kFull, nn, off);
        mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, off));
      }
      if (lane == 0) {
        // fp16 shadow: the query is scaled by 2^f so that its largest magnitude lands in [2^13, 2^14) (as the rows are,
        // synth.cu shadow_convert_kernel); the scan multiplies the accumulators by 2^-f
        int f = 0;
        if (BF16 && (flags & kFlagShadowScan) && mx > 0.f && mx <= 3.0e38f) { int ex; frexpf(mx, &ex); f = 14 - ex; }
        s_up1 = exp2f(float(f / 2)); s_up2 = exp2f(float(f - f / 2));
        if (qscale && live) qscale[r] = exp2f(float(-f));
        s_inv = inv;
        if (live) {
          qinv[r] = inv;
          qerr[r] = tc_filter_relerr(dim, BF16, BF16 && (flags & kFlagShadowScan)) * sqrtf(nn) *
                    ((flags & DEVQA_DEBUG_TIGHT_BOUND) ? 0.015625f : 1.f);
        }
      }
    }
    __syncthreads();
    const float inv = s_inv;
    const int pass = r / qp, rp = r % qp;
    if constexpr (BF16) {
      __nv_bfloat16* base = static_cast<__nv_bfloat16*>(qcat_) + size_t(pass) * n_kblocks * (qp * 64);
      if (flags & kFlagShadowScan) {                     // scaled fp16 (same 2-byte slots)
        const float up1 = s_up1, up2 = s_up2;
        __half* hbase = reinterpret_cast<__half*>(base);
        for (int c = tid; c < d_pad; c += nthr) {
          const int kb = c / 64, w = c % 64;
          hbase[size_t(kb) * (qp * 64) + rp * 64 + (((w >> 3) ^ (rp & 7)) << 3) + (w & 7)] = __float2half_rn(srow[c] * inv * up1 * up2);
        }
      } else {
        for (int c = tid; c < d_pad; c += nthr) {
          const int kb = c / 64, w = c % 64;
          base[size_t(kb) * (qp * 64) + rp * 64 + (((w >> 3) ^ (rp & 7)) << 3) + (w & 7)] = __float2bfloat16_rn(srow[c] * inv);
        }
      }
    } else {
      float* base = static_cast<float*>(qcat_) + size_t(pass) * n_kblocks * (qp * kBlockK);
      for (int c4 = tid; c4 < d_pad / 4; c4 += nthr) {                  // one 16-byte chunk of the swizzled row per store
        const float4 v = reinterpret_cast<const float4*>(srow)[c4];
        uint32_t h0, h1, h2, h3;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h0) : "f"(v.x * inv));
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h1) : "f"(v.y * inv));
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h2) : "f"(v.z * inv));
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h3) : "f"(v.w * inv));
        const int kb = c4 / (kBlockK / 4), w4 = c4 % (kBlockK / 4);
        float4* dst = reinterpret_cast<float4*>(base + size_t(kb) * (qp * kBlockK) + rp * kBlockK) + (w4 ^ (rp & 7));
        *dst = make_float4(__uint_as_float(h0), __uint_as_float(h1), __uint_as_float(h2), __uint_as_float(h3));
      }
    }
  }
}

// bf16-bank variant: three bf16 planes hi + mid + lo (24 mantissa bits together), k-blocks of 64 columns.
// Output per pass: [n_kblocks][3*qp rows][64 bf16], rows 0..qp-1 = hi, qp.. = mid, 2qp.. = lo, 128B-swizzled.
__global__ void __launch_bounds__(256) prep_queries_bf16_kernel(const float* __restrict__ q, int64_t q_ld, int n_queries, int dim,
                                                                int qp, int n_kblocks, int q_pad, unsigned flags,
                                                                __nv_bfloat16* __restrict__ qcat, uint32_t* __restrict__ gthr,
                                                                uint32_t* __restrict__ gmax, int gmax_stride) {
  pdl_wait();
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < q_pad; i += blockDim.x) gthr[i] = 0u;
  if (gmax)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < q_pad * gmax_stride; i += gridDim.x * blockDim.x) gmax[i] = 0u;
  const int lane = threadIdx.x & 31;
  const int warps = gridDim.x * (blockDim.x >> 5);
  const int d_pad = n_kblocks * 64;
  for (int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < q_pad; r += warps) {
    float inv = 0.f;
    const float* src = q + size_t(r < n_queries ? r : 0) * q_ld;
    if (r < n_queries) {
      inv = 1.f;
      if (flags & DEVQA_NORMALIZE_QUERY) {
        float ss = 0.f;
        for (int c = lane; c < dim; c += 32) { const float x = src[c]; ss = fmaf(x, x, ss); }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) ss += __shfl_xor_sync(kFull, ss, off);
        const float nrm = sqrtf(ss);
        inv = 1.0f / ((flags & DEVQA_NOEPS) ? nrm : fmaxf(nrm, kNormEps));
      }
    }
    const int pass = r / qp, rp = r % qp;
    __nv_bfloat16* base = qcat + size_t(pass) * n_kblocks * (3 * qp * 64);
    for (int c = lane; c < d_pad; c += 32) {
      const float x = (r < n_queries && c < dim) ? src[c] * inv : 0.f;
      const __nv_bfloat16 hi = __float2bfloat16_rn(x);
      const float r1 = x - __bfloat162float(hi);
      const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
      const __nv_bfloat16 lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
      const int kb = c / 64, w = c % 64;
      __nv_bfloat16* blk = base + size_t(kb) * (3 * qp * 64);
      const __nv_bfloat16 part[3] = {hi, mid, lo};
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        const int row = t * qp + rp;
        blk[row * 64 + (((w >> 3) ^ (row & 7)) << 3) + (w & 7)] = part[t];
      }
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
      cudaGetLastError();
      set_error("cuTensorMapEncodeTiled is not available from the driver");
      return nullptr;
    }
    encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  return encode;
}

int make_map(CUtensorMap* map, const void* base, bool bf16, uint64_t cols, uint64_t rows, uint64_t ld_bytes, uint32_t box_rows,
             bool f16 = false) {
  const EncodeTiledFn encode = tensor_map_encoder();
  if (!encode) return DEVQA_E_CUDA;
  const cuuint64_t gdim[2] = {cols, rows};
  const cuuint64_t gstride[1] = {ld_bytes};
  const cuuint32_t box[2] = {bf16 ? 64u : 32u, box_rows};     // 128 bytes of columns either way
  const cuuint32_t estride[2] = {1, 1};
  CUresult r = encode(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estride,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (cols %llu rows %llu ld %llu)", int(r),
              (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)ld_bytes);
    return DEVQA_E_CUDA;
  }
  return DEVQA_OK;
}

// [rows][128 bytes], no swizzle (the query operand is stored pre-swizzled): box = box_rows whole rows
int make_rows_map(CUtensorMap* map, const void* base, uint64_t rows, uint32_t box_rows) {
  const EncodeTiledFn encode = tensor_map_encoder();
  if (!encode) return DEVQA_E_CUDA;
  const cuuint64_t gdim[2] = {128, rows ? rows : 1};
  const cuuint64_t gstride[1] = {128};
  const cuuint32_t box[2] = {128, box_rows};
  const cuuint32_t estride[2] = {1, 1};
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), gdim, gstride, box, estride,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (query operand) failed with CUresult %d (rows %llu)", int(r), (unsigned long long)rows);
    return DEVQA_E_CUDA;
  }
  return DEVQA_OK;
}

}  // namespace

bool scan_tc_supported(int bank_dtype, const void* bank, int64_t dim, int64_t bank_ld, int top_k) {
  const int64_t align = bank_dtype == DEVQA_F32 ? 4 : 8;      // 16-byte rows for the tensor map
  return reinterpret_cast<uintptr_t>(bank) % 16 == 0 && bank_ld % align == 0 && dim >= 1 && topk_capacity(top_k) <= 128;
}

// bank columns per 128-byte k-block row, and the padded k-block count
static int elems_per_kblock(int mode) { return (mode == kTcBf16x3 || mode == kTcBf16x1) ? 64 : 32; }
static int mode_parts(int mode) { return mode == kTcBf16x3 ? 3 : ((mode == kTcTf32x1 || mode == kTcBf16x1) ? 1 : 2); }
int scan_tc_kblocks(int mode, int64_t dim) { return int((dim + elems_per_kblock(mode) - 1) / elems_per_kblock(mode)); }
// bytes of the prepared query operand for q_pad queries
size_t scan_tc_qcat_bytes(int mode, int64_t dim, int q_pad) {
  return size_t(q_pad) * scan_tc_kblocks(mode, dim) * 128 * mode_parts(mode);
}

// Queries per pass + pipeline depth that fit in shared memory / tensor memory for this list capacity.
int scan_tc_plan(int mode, int n_queries, int64_t dim, int kcap, bool cached_norms, int64_t n_tiles, int sm_count,
                 unsigned flags, int* qp_out, int* stages_out, bool* pair_out, bool oneshot, bool emit, bool* resident_out) {
  *pair_out = false;
  if (resident_out) *resident_out = false;
  const bool filt = mode == kTcTf32x1 || mode == kTcBf16x1;
  // more than one K-set (1024 fp32 / 2048 bf16 columns) sums K-sets in the epilogue's registers (MULTI):
  // 32 queries per pass at most.  The filter keeps the whole K chain in the tensor core (its error bound
  // covers the drift) and, with cached row norms (no CUDA-core pass over the tile), takes 128 queries per pass.
  int qp = filt ? (cached_norms ? 256 : 64) : (scan_tc_kblocks(mode, dim) > kSetKBlocks ? 32 : 64);
  const int q16 = (n_queries + 15) / 16 * 16;
  while (qp > 16 && qp / 2 >= q16) qp >>= 1;
  // banks of fewer tiles than SMs: narrower passes run side by side, so that tiles x passes covers the machine
  if (filt && oneshot) {
    // (tried: ONE 256-query pass on CTA pairs, half of the query operand per CTA -- 185 us against 150 at C2: 64 SMs
    // cannot pull the operands out of L2 as fast as 128 can)
    // one tile per CTA, every CTA of the launch co-resident: halve the pass width only while one launch still holds
    // all passes (C2: 64 tiles, 256 queries -> two 128-query passes = 128 CTAs; four 64-query passes would not fit)
    while (qp > 32 && n_tiles * ((n_queries + qp / 2 - 1) / (qp / 2)) <= sm_count) qp >>= 1;
  } else if (filt) {
    while (qp > 32 && n_tiles * ((n_queries + qp - 1) / qp) < sm_count) qp >>= 1;
  }
  for (; qp >= 16; qp >>= 1) {
    // 256 queries per pass on a long scan: CTA pairs (each CTA holds half of the query operand)
    const bool pair = filt && cached_norms && (qp == 256 || (qp >= 64 && (flags & kDbgPairAll))) &&
                      n_tiles >= 2 * int64_t(sm_count) && sm_count % 2 == 0 && !(flags & kDbgNoPair);
    const int lcap = (filt && (oneshot || emit)) ? 0 : kcap;
    // long filter scans: the whole query operand resident, if at least 5 ring stages of bank tiles still fit
    // (What a wide pass can stream is bytes in flight / ring latency -- ~46 B/clk and SM with ~190 KB in flight, DESIGN.md
    // 3.1h: the query operand out of the ring leaves it all to bank tiles.  A CTA pair holds half of the operand per
    // CTA; 128-query passes AS PAIRS just to make it fit were tried: +4.5 % on a 1.25M-row shard, -4 % at 10M rows.)
    if (filt && emit && resident_out && !(flags & kDbgNoResident) && 2 * qp <= 512) {
      const int nkb = scan_tc_kblocks(mode, dim);
      for (int stages = kMaxStages; stages >= 5; --stages) {
        if (tc_layout(stages, qp, lcap, mode_parts(mode), pair, nkb).total + 1024 <= 225 * 1024) {
          *qp_out = qp; *stages_out = stages; *resident_out = true; *pair_out = pair;
          return DEVQA_OK;
        }
      }
    }
    for (int stages = kMaxStages; stages >= 3; --stages) {
      // TMEM: 2 accumulator buffers of 3*qp (filter: qp) columns (+ one 32-column lo tile per stage for 3xTF32)
      if ((filt ? 2 : 6) * qp + (mode == kTcTf32x3 ? stages * kBlockK : 0) > 512) continue;
      if (tc_layout(stages, qp, lcap, mode_parts(mode), pair).total + 1024 <= 225 * 1024) {
        *qp_out = qp; *stages_out = stages; *pair_out = pair;
        return DEVQA_OK;
      }
    }
  }
  set_error("tcgen05 scan: no shared-memory plan for kcap %d", kcap);
  return DEVQA_E_UNSUPPORTED;
}

// qcat must hold scan_tc_qcat_bytes(); q_pad is a multiple of qp; gthr [q_pad] is zeroed here.
int scan_tc_prep_queries(int mode, const float* queries, int64_t query_ld, int n_queries, int qp, int q_pad,
                         int64_t dim, unsigned flags, void* qcat, uint32_t* gthr, float* qinv, float* qerr,
                         uint32_t* seed_u, uint32_t* gmax, int gmax_stride, const uint32_t* need, cudaStream_t stream,
                         uint32_t* zero_words, int n_zero, float* qscale) {
  const int n_kblocks = scan_tc_kblocks(mode, dim);
  const int grid = (q_pad + 7) / 8;
  if (mode == kTcTf32x3)
    DEVQA_CUDA(launch_pdl(prep_queries_kernel, dim3(grid), dim3(256), 0, stream, queries, query_ld, n_queries, int(dim), qp,
                          n_kblocks, q_pad, flags, static_cast<float*>(qcat), gthr, need, gmax, gmax_stride));
  else if (mode == kTcBf16x3)
    DEVQA_CUDA(launch_pdl(prep_queries_bf16_kernel, dim3(grid), dim3(256), 0, stream, queries, query_ld, n_queries, int(dim),
                          qp, n_kblocks, q_pad, flags, static_cast<__nv_bfloat16*>(qcat), gthr, gmax, gmax_stride));
  else {
    const int ek = mode == kTcBf16x1 ? 64 : 32;
    const size_t smem = size_t(n_kblocks) * ek * 4;
    const int fgrid = q_pad < 1184 ? q_pad : 1184;          // one CTA per query (8 per SM at most)
    auto kern = mode == kTcBf16x1 ? prep_queries_filt_kernel<true> : prep_queries_filt_kernel<false>;
    if (smem > 40 * 1024) DEVQA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    DEVQA_CUDA(launch_pdl(kern, dim3(fgrid), dim3(128), smem, stream, queries, query_ld, n_queries, int(dim), qp, n_kblocks,
                          q_pad, flags, qcat, gthr, qinv, qerr, seed_u, gmax, gmax_stride, zero_words, n_zero, qscale));
  }
  DEVQA_CUDA(cudaGetLastError());
  count_launch();
  return DEVQA_OK;
}

// Does this mode/dim run the scan-and-queue epilogue with the cross-CTA threshold exchange (needs gmax, no seeding)?
bool scan_tc_exchanges_thresholds(int mode, int64_t dim) { return mode == kTcTf32x1 || mode == kTcBf16x1 || scan_tc_kblocks(mode, dim) <= kSetKBlocks; }

static long long* g_tc_trace = nullptr;
void scan_tc_set_trace(long long* buf) { g_tc_trace = buf; }

int scan_tc_launch(const ScanParams& sp, int mode, const void* qcat, uint32_t* gthr, int qp, int stages, int passes,
                   int ctas_per_pass, bool pair, cudaStream_t stream) {
  const int grid = passes * ctas_per_pass;
  const bool bf = mode == kTcBf16x3 || mode == kTcBf16x1;
  const bool filt = mode == kTcTf32x1 || mode == kTcBf16x1;
  const int parts = mode_parts(mode);
  const int n_kblocks = scan_tc_kblocks(mode, sp.dim);
  static const bool hint_set = [] {
    const char* e = getenv("DEVQA_WAIT_HINT_NS");
    const uint32_t ns = e ? uint32_t(strtoul(e, nullptr, 10)) : 0u;
    if (ns) cudaMemcpyToSymbol(c_wait_hint_ns, &ns, sizeof(ns));
    return true;
  }();
  (void)hint_set;
  CUtensorMap bank_map;
  int rc;
  if ((sp.flags & kDbgContigTiles) && sp.bank_ld == sp.dim && sp.dim % (bf ? 64 : 32) == 0)
    // timing experiment: the same bytes addressed as [rows * k-blocks][128 bytes] -- every box one contiguous 16 KB run
    rc = make_map(&bank_map, sp.bank, bf, bf ? 64 : 32, uint64_t(sp.n_rows) * uint64_t(n_kblocks), 128, kTileRows,
                  bf && (sp.flags & kFlagShadowScan) != 0);
  else
    rc = make_map(&bank_map, sp.bank, bf, uint64_t(sp.dim), uint64_t(sp.n_rows), uint64_t(sp.bank_ld) * (bf ? 2 : 4), kTileRows,
                  bf && (sp.flags & kFlagShadowScan) != 0);
  if (rc) return rc;
  TcParams p{};
  p.n_rows = sp.n_rows; p.dim = int(sp.dim); p.n_kblocks = n_kblocks;
  p.n_tiles = int((sp.n_rows + kTileRows - 1) / kTileRows);
  p.qp = qp; p.nq_pass = sp.nq_pass; p.q0 = sp.q0; p.n_queries = sp.n_queries;
  p.top_k = sp.top_k; p.kcap = sp.kcap; p.stages = stages; p.passes = passes; p.flags = sp.flags;
  p.cand = sp.cand; p.gthr = gthr; p.idmap = sp.idmap; p.row_sumsq = sp.row_sumsq;
  p.row_scale = sp.row_scale; p.qscale = sp.qscale;
  p.need = sp.need;
  p.oneshot = sp.oneshot; p.emit = sp.emit; p.gbar = sp.gbar; p.ocount = sp.ocount; p.ocand = sp.ocand; p.ocap = sp.ocap;
  p.gmax = sp.gmax; p.gmax_stride = sp.gmax_stride; p.lostu = sp.lostu; p.qerr = sp.qerr; p.seed_u = sp.seed_u; p.bank_max_sumsq = sp.bank_max_sumsq; p.k_true = sp.k_true > 0 ? sp.k_true : sp.top_k;
  // pass index = q0 / qp; one pass = n_kblocks * parts * qp rows of 128 bytes (32 words)
  p.qcat = static_cast<const float*>(qcat) + size_t(sp.q0 / qp) * size_t(n_kblocks) * parts * qp * 32;
  p.trace = g_tc_trace;
  if (pair) {
    // the query operand as a matrix of 128-byte rows: [passes x k-blocks x qp]; each CTA of a pair loads qp/2 rows per stage
    const uint64_t q_rows = uint64_t((sp.n_queries - sp.q0 + qp - 1) / qp) * uint64_t(n_kblocks) * uint64_t(parts * qp);
    if ((rc = make_rows_map(&p.qmap, p.qcat, q_rows, uint32_t(qp / 2)))) return rc;
  }
  p.b_resident = (filt && sp.b_resident) ? 1 : 0;
  const size_t smem = tc_layout(stages, qp, (filt && (sp.oneshot || sp.emit)) ? 0 : sp.kcap, parts, pair, p.b_resident ? n_kblocks : 0).total + 1024;
  const bool multi = !filt && p.n_kblocks > kSetKBlocks;
  void (*kern)(CUtensorMap, TcParams) = nullptr;
  if (mode == kTcTf32x3) {
    if (!multi) kern = qp == 64 ? scan_tc_kernel<64, false, kTf32x3> : qp == 32 ? scan_tc_kernel<32, false, kTf32x3> : scan_tc_kernel<16, false, kTf32x3>;
    else kern = qp == 32 ? scan_tc_kernel<32, true, kTf32x3> : scan_tc_kernel<16, true, kTf32x3>;
  } else if (mode == kTcBf16x3) {
    if (!multi) kern = qp == 64 ? scan_tc_kernel<64, false, kBf16x3> : qp == 32 ? scan_tc_kernel<32, false, kBf16x3> : scan_tc_kernel<16, false, kBf16x3>;
    else kern = qp == 32 ? scan_tc_kernel<32, true, kBf16x3> : scan_tc_kernel<16, true, kBf16x3>;
  } else if (mode == kTcBf16x1) {
    kern = qp == 256 ? scan_tc_kernel<256, false, kBf16x1> : qp == 128 ? scan_tc_kernel<128, false, kBf16x1> : qp == 64 ? scan_tc_kernel<64, false, kBf16x1>
         : qp == 32 ? scan_tc_kernel<32, false, kBf16x1> : scan_tc_kernel<16, false, kBf16x1>;
  } else {
    kern = qp == 256 ? scan_tc_kernel<256, false, kTf32x1> : qp == 128 ? scan_tc_kernel<128, false, kTf32x1> : qp == 64 ? scan_tc_kernel<64, false, kTf32x1>
         : qp == 32 ? scan_tc_kernel<32, false, kTf32x1> : scan_tc_kernel<16, false, kTf32x1>;
  }
  const bool qp_ok = qp == 16 || qp == 32 || (qp == 64 && !multi) || ((qp == 128 || qp == 256) && filt);
  if (!qp_ok) {
    set_error("tcgen05 scan: unsupported queries-per-pass %d", qp);
    return DEVQA_E_INVALID;
  }
  if (pair) {
    DEVQA_REQUIRE(filt && qp >= 64 && passes == 1 && grid % 2 == 0, "tcgen05 scan: bad CTA-pair launch");
    if (mode == kTcBf16x1) {
      kern = qp == 256 ? scan_tc_kernel<256, false, kBf16x1, true> : qp == 128 ? scan_tc_kernel<128, false, kBf16x1, true>
                                                                                : scan_tc_kernel<64, false, kBf16x1, true>;
    } else {
      kern = qp == 256 ? scan_tc_kernel<256, false, kTf32x1, true> : qp == 128 ? scan_tc_kernel<128, false, kTf32x1, true>
                                                                                : scan_tc_kernel<64, false, kTf32x1, true>;
    }
  }
  DEVQA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  if (pair) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(unsigned(grid)); cfg.blockDim = dim3(kTcThreads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
    DEVQA_CUDA(cudaLaunchKernelEx(&cfg, kern, bank_map, p));
  } else if (p.oneshot) {
    // the CTAs wait for each other (oneshot_epilogue): a cooperative launch guarantees they are all resident
    DEVQA_REQUIRE(filt && ctas_per_pass == p.n_tiles, "tcgen05 scan: one-shot mode needs one CTA per tile");
    // DEVQA_ONESHOT_LAUNCH (experiments): "coop" = cooperative only, "pdl" = programmatic dependent launch only (grid <=
    // SM count and one CTA per SM: resident in practice, and grid_wait gives up after a second), default = both.
    static const char* how = getenv("DEVQA_ONESHOT_LAUNCH");
    const bool want_coop = !how || strcmp(how, "pdl") != 0;
    const bool want_pdl = pdl_enabled() && (!how || strcmp(how, "coop") != 0);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(unsigned(grid)); cfg.blockDim = dim3(kTcThreads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (want_coop) { attr[na].id = cudaLaunchAttributeCooperative; attr[na].val.cooperative = 1; ++na; }
    if (want_pdl) { attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[na].val.programmaticStreamSerializationAllowed = 1; ++na; }
    cfg.attrs = attr; cfg.numAttrs = na;
    cudaError_t le = cudaLaunchKernelEx(&cfg, kern, bank_map, p);
    if (le != cudaSuccess && na == 2) {                      // the combination is not accepted: cooperative alone
      cudaGetLastError();
      cfg.numAttrs = 1;
      le = cudaLaunchKernelEx(&cfg, kern, bank_map, p);
      static bool told = false;
      if (!told && getenv("DEVQA_VERBOSE")) { told = true; fprintf(stderr, "libdevqa: cooperative + PDL launch rejected, using cooperative\n"); }
    }
    DEVQA_CUDA(le);
  } else {
    DEVQA_CUDA(launch_pdl(kern, dim3(unsigned(grid)), dim3(kTcThreads), smem, stream, bank_map, p));
  }
  DEVQA_CUDA(cudaGetLastError());
  count_launch();
  return DEVQA_OK;
}

}  // namespace devqa
