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
};

size_t scan_stream_smem_bytes(int64_t dim, int nv, int qt, int kcap);
int scan_stream_rows_per_group(int qt);
int scan_stream_launch(const ScanParams& p, int bank_dtype, bool vectorised, int qt, int grid,
                       cudaStream_t stream);

// scan_tc.cu: tcgen05 / TMEM / TMA contraction kernel for large query batches (fp32 and bf16 banks)
bool scan_tc_supported(int bank_dtype, const void* bank, int64_t dim, int64_t bank_ld, int top_k);
int scan_tc_plan(int bank_dtype, int n_queries, int64_t dim, int kcap, int* qp_out, int* stages_out);
size_t scan_tc_qcat_bytes(int bank_dtype, int64_t dim, int q_pad);
// queries -> normalised hi/lo (TF32, fp32 banks) or hi/mid/lo (bf16 banks) split, laid out per pass as the
// MMA B operand (see scan_tc.cu); also zeroes the shared threshold array gthr [q_pad]
int scan_tc_prep_queries(int bank_dtype, const float* queries, int64_t query_ld, int n_queries, int qp, int q_pad,
                         int64_t dim, unsigned flags, void* qcat, uint32_t* gthr, cudaStream_t stream);
// one launch = `passes` consecutive query passes starting at p.q0, each on ctas_per_pass CTAs
int scan_tc_launch(const ScanParams& p, int bank_dtype, const void* qcat, uint32_t* gthr, int qp, int stages, int passes,
                   int ctas_per_pass, cudaStream_t stream);
void scan_tc_set_trace(long long* buf);

// merge.cu: reduce `n_lists` sorted lists per query ([n_lists, n_queries, list_stride], the first
// list_len entries of each valid) to the final top_k.  Exactly one of (out_keys) or
// (out_scores, out_ids) is written: [n_queries, top_k].  With kth_codes != nullptr nothing else is written:
// kth_codes[q] receives the monotone score code of the k-th best key (0 when fewer than k exist).
int merge_lists_launch(const uint64_t* lists, int n_lists, int64_t n_queries, int list_stride,
                       int list_len, int top_k, uint64_t* out_keys, float* out_scores,
                       int64_t* out_ids, cudaStream_t stream, uint32_t* kth_codes = nullptr);

// synth.cu
int fill_synthetic_launch(void* dst, int dtype, int64_t n_rows, int64_t dim, int64_t ld, uint64_t seed,
                          int64_t first_global_row, int kind, cudaStream_t stream);
// norm.cu
int normalize_rows_launch(const float* src, int64_t src_ld, float* dst, int64_t dst_ld, int64_t n_rows, int64_t dim,
                          unsigned flags, cudaStream_t stream);
// norm.cu: ||row||^2 of n rows (fp32 or bf16 storage), one warp per row
int row_sumsq_launch(const void* rows, int dtype, int64_t n_rows, int64_t dim, int64_t ld, float* out, cudaStream_t stream);
// synth.cu helper kernels
int widen_rows_launch(const void* src, int64_t src_ld, float* dst, int64_t dst_ld, int64_t n_rows, int64_t dim,
                      cudaStream_t stream);
int convert_rows_launch(const float* src, int64_t src_ld, void* dst, int dst_dtype, int64_t dst_ld,
                        int64_t n_rows, int64_t dim, cudaStream_t stream);

}  // namespace devqa
