// Internal interface between the C ABI (api.cu) and the kernels.
#pragma once
#include "common.cuh"

namespace devqa {

constexpr int kScanThreads = 512;   // 16 warps, one CTA per SM (persistent grid)

struct ScanParams {
  const void* bank;        // [n_rows, dim], row stride bank_ld elements
  int64_t n_rows;
  int64_t dim;
  int64_t bank_ld;
  const float* queries;    // [n_queries, dim], row stride query_ld
  int64_t query_ld;
  int32_t n_queries;       // total queries of the call (candidate buffer layout)
  int32_t q0;              // first query of this pass
  int32_t nq_pass;         // queries in this pass (<= QT)
  int32_t top_k;
  int32_t kcap;            // topk_capacity(top_k)
  uint32_t flags;
  uint64_t* cand;          // [grid, n_queries, kcap] sorted candidate lists (output)
  IdMap idmap;
  const float* row_sumsq;  // optional cached ||row||^2 of library-owned banks (nullptr: computed in-kernel)
  const uint32_t* need;    // tcgen05 scan as a conditional fallback: [n_queries] flags, passes with none set exit at once
  // tcgen05 filter variant
  const float* qerr;       // [n_queries] tc_filter_relerr(dim) * ||query as used||
  const uint32_t* seed_u;  // optional [n_queries] score codes of a prefix sample's k-th best upper bound
  const uint32_t* bank_max_sumsq;   // un-normalised scores: max ||row||^2 over the bank (float bits)
  int32_t k_true;          // caller's top_k (the lists keep kcap entries)
  // stream kernel, capacity 32: the last CTA to finish merges the per-CTA lists itself (no merge launch)
  uint32_t* ticket;        // zeroed device word (re-zeroed by the kernel); nullptr: separate merge kernel
  uint64_t* out_keys; float* out_scores; int64_t* out_ids;   // final [n_queries, top_k] (exactly one form)
  uint32_t* lostu;         // [grid, n_queries] out: bound on the rows each CTA's full list lost (score code, 0 = none)
  uint32_t* gmax;          // [q_pad, gmax_stride] zeroed scratch: each CTA's best upper bound per query (score codes)
  int32_t gmax_stride;
};

size_t scan_stream_smem_bytes(int64_t dim, int nv, int qt, int kcap);
int scan_stream_rows_per_group(int qt);
int scan_stream_launch(const ScanParams& p, int bank_dtype, bool vectorised, int qt, int grid,
                       cudaStream_t stream);

// scan_tc.cu: tcgen05 / TMEM / TMA contraction kernel for large query batches.  `mode` selects the arithmetic:
//   kTcTf32x3 (= DEVQA_F32)  fp32 bank, fp32-exact scores (3xTF32 split)
//   kTcBf16x3 (= DEVQA_BF16) bf16 bank, exact w.r.t. the rounded bank
//   kTcTf32x1                fp32 bank, approximate scores with a proven error bound: a FILTER whose top-K'
//                            lists are rescored exactly by refine.cu
constexpr int kTcTf32x3 = 0, kTcBf16x3 = 1, kTcTf32x1 = 2, kTcBf16x1 = 3;   // kTcBf16x1: the filter on a bf16 bank
// |approximate - exact| <= tc_filter_relerr(dim) * ||row|| * ||query||  for the kTcTf32x1 scores:
//   2^-10 (bank words truncated to TF32) + 2^-11 (queries rounded to TF32) + their product, by Cauchy-Schwarz;
//   dim * 2^-24 for the accumulation: the tensor core adds dim/8 times into an fp32 accumulator with truncation
//   (<= 2 ulp of the running magnitude each: dim * 2^-25) and the fp32 re-scoring it is compared with rounds
//   (dim/32 + 5) times to nearest.  tests/test_gpu_filter.py measures the actual error against this bound.
// bf16 banks (kTcBf16x1): the bank is exact, queries are rounded to bf16 (<= 2^-9): 1.01 * 2^-9 + dim * 2^-24.
__host__ __device__ inline float tc_filter_relerr(int64_t dim, bool bf16_bank = false) {
  return (bf16_bank ? 1.01f * 0.001953125f : 1.5f * 0.0009765625f) + float(dim) * 5.9604644775390625e-08f;
}
bool scan_tc_supported(int bank_dtype, const void* bank, int64_t dim, int64_t bank_ld, int top_k);
int scan_tc_plan(int mode, int n_queries, int64_t dim, int kcap, bool cached_norms, int64_t n_tiles, int sm_count,
                 unsigned flags, int* qp_out, int* stages_out, bool* pair_out);
size_t scan_tc_qcat_bytes(int mode, int64_t dim, int q_pad);
bool scan_tc_exchanges_thresholds(int mode, int64_t dim);   // scan-and-queue epilogue + gmax exchange (else: seeded score-tile epilogue)
// queries -> normalised, split (hi/lo TF32; hi/mid/lo bf16; hi TF32 only for the filter), laid out per pass as the
// MMA B operand (see scan_tc.cu); also zeroes the shared threshold array gthr [q_pad].  Filter mode also writes,
// per query, the scale the normalisation applied (qinv) and the error scale qerr = tc_filter_relerr(dim) *
// ||query as used||, and zeroes seed_u [q_pad] and the threshold exchange table gmax [q_pad, gmax_stride].
int scan_tc_prep_queries(int mode, const float* queries, int64_t query_ld, int n_queries, int qp, int q_pad,
                         int64_t dim, unsigned flags, void* qcat, uint32_t* gthr, float* qinv, float* qerr,
                         uint32_t* seed_u, uint32_t* gmax, int gmax_stride, const uint32_t* need, cudaStream_t stream);
// one launch = `passes` consecutive query passes starting at p.q0, each on ctas_per_pass CTAs
// (pair: clusters of two CTAs sharing the query operand, see scan_tc.cu)
int scan_tc_launch(const ScanParams& p, int mode, const void* qcat, uint32_t* gthr, int qp, int stages, int passes,
                   int ctas_per_pass, bool pair, cudaStream_t stream);
void scan_tc_set_trace(long long* buf);

// refine.cu: exact fp32 re-scoring of the filter's candidates (bit-identical to scan_stream.cu's arithmetic).
struct RefineParams {
  const void* bank; int bank_dtype; int64_t n_rows, dim, bank_ld;   // fp32 or bf16 rows (DEVQA_F32 / DEVQA_BF16)
  const float* queries; int64_t query_ld; int32_t n_queries;
  const float* qinv;          // [n_queries] query scale applied by the normalisation (1 when off)
  const uint64_t* akeys;      // [n_queries, kp] merged candidates: keys on the upper bound u, sorted descending, 0 = empty
  const uint32_t* lostu;      // [n_lists, n_queries] per-CTA bound on the rows its full list lost (score code, 0 = none)
  int32_t n_lists;
  // refine_lists_launch only: the per-CTA lists themselves and what the CTAs exchanged
  const uint64_t* cand;       // [n_lists, n_queries, cand_kcap]
  int32_t cand_kcap;
  const uint32_t* gmax;       // [n_queries (padded), gmax_stride] each CTA's best upper bound per query (score codes)
  int32_t gmax_stride;
  const float* qerr;          // [n_queries] tc_filter_relerr(dim) * ||query as used||
  const uint32_t* bank_max_sumsq;   // un-normalised scores: max ||row||^2 (float bits)
  int32_t kp, top_k;
  uint32_t flags;
  IdMap idmap;
  const uint32_t* only_if;    // refine_launch only: when set, re-score (and re-sort) akeys of the flagged queries and
                              // nothing else -- no completeness verdict (used on the exact fallback's result)
  uint32_t* need;             // [n_queries] out: 1 when a lost row could still belong to the exact top-k
  uint32_t* fallbacks;        // device counter of such queries (devqa_filter_fallbacks)
  uint64_t* out_keys; float* out_scores; int64_t* out_ids;   // exactly one of keys / (scores, ids): [n_queries, top_k]
};
int refine_launch(const RefineParams& p, cudaStream_t stream);          // from the merged list akeys
int refine_lists_launch(const RefineParams& p, cudaStream_t stream);    // from the per-CTA lists (n_lists >= top_k)

// merge.cu: reduce `n_lists` sorted lists per query ([n_lists, n_queries, list_stride], the first
// list_len entries of each valid) to the final top_k.  Exactly one of (out_keys) or
// (out_scores, out_ids) is written: [n_queries, top_k].  With kth_codes != nullptr nothing else is written:
// kth_codes[q] receives the monotone score code of the k-th best key (0 when fewer than k exist).
// With need != nullptr only queries whose flag is set are touched.
int merge_lists_launch(const uint64_t* lists, int n_lists, int64_t n_queries, int list_stride,
                       int list_len, int top_k, uint64_t* out_keys, float* out_scores,
                       int64_t* out_ids, cudaStream_t stream, uint32_t* kth_codes = nullptr,
                       const uint32_t* need = nullptr);

// synth.cu
int fill_synthetic_launch(void* dst, int dtype, int64_t n_rows, int64_t dim, int64_t ld, uint64_t seed,
                          int64_t first_global_row, int kind, cudaStream_t stream);
// norm.cu
int normalize_rows_launch(const float* src, int64_t src_ld, float* dst, int64_t dst_ld, int64_t n_rows, int64_t dim,
                          unsigned flags, cudaStream_t stream);
// norm.cu: ||row||^2 of n rows (fp32 or bf16 storage), one warp per row
// (max_bits: optional device word that receives atomicMax of the sums' bit patterns)
int row_sumsq_launch(const void* rows, int dtype, int64_t n_rows, int64_t dim, int64_t ld, float* out, cudaStream_t stream,
                     uint32_t* max_bits = nullptr);
// synth.cu helper kernels
int widen_rows_launch(const void* src, int64_t src_ld, float* dst, int64_t dst_ld, int64_t n_rows, int64_t dim,
                      cudaStream_t stream);
int convert_rows_launch(const float* src, int64_t src_ld, void* dst, int dst_dtype, int64_t dst_ld,
                        int64_t n_rows, int64_t dim, cudaStream_t stream);

}  // namespace devqa
