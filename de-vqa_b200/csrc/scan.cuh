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
  const float* row_scale;  // fp16 shadow: 2^-e per row (the accumulators of row r are multiplied by it); else nullptr
  const float* qscale;     // fp16 shadow: 2^-f per query (written by the query-prep kernel); else nullptr
  const uint32_t* need;    // tcgen05 scan as a conditional fallback: [n_queries] flags, passes with none set exit at once
  // tcgen05 filter variant
  const float* qerr;       // [n_queries] tc_filter_relerr(dim) * ||query as used||
  const uint32_t* seed_u;  // optional [n_queries] score codes of a prefix sample's k-th best upper bound
  const uint32_t* bank_max_sumsq;   // un-normalised scores: max ||row||^2 over the bank (float bits)
  int32_t k_true;          // caller's top_k (the lists keep kcap entries)
  // stream kernel, capacity 32: the last CTA to finish merges the per-CTA lists itself (no merge launch)
  uint32_t* ticket;        // zeroed device word (re-zeroed by the kernel); nullptr: separate merge kernel
  uint64_t* out_keys; float* out_scores; int64_t* out_ids;   // final [n_queries, top_k] (exactly one form)
  // stream kernel, deep top-k (top_k > DEVQA_MAX_TOPK, one pass per 1024 results): only keys BELOW below[q] take part
  const uint64_t* below;   // [n_queries] exclusive upper bound keys (0 = none); nullptr = ordinary scan
  // stream kernel, threshold-select mode (select.cu): mask[q][row] = score * sel_scale > sel_thr[q]
  const float* sel_thr;    // [n_queries] device thresholds
  float sel_scale;
  uint8_t* sel_mask;       // [n_queries, sel_mask_ld] out; nullptr = ordinary top-k scan
  int64_t sel_mask_ld;
  uint32_t* lostu;         // [grid, n_queries] out: bound on the rows each CTA's full list lost (score code, 0 = none)
  uint32_t* gmax;          // [q_pad, gmax_stride] zeroed scratch: each CTA's best upper bound per query (score codes)
  int32_t gmax_stride;
  // tcgen05 filter, one-shot mode (every CTA owns one tile; scan_tc.cu oneshot_epilogue)
  int32_t oneshot;
  int32_t emit;            // long scans: candidates to ocand instead of per-CTA lists (scan_tc.cu emit_epilogue)
  int32_t b_resident;      // scan_tc_plan's choice: the pass's query operand stays in shared memory
  uint32_t* gbar;          // [2 * passes] zeroed arrival counters
  uint32_t* ocount;        // [n_queries] zeroed candidate counts
  uint64_t* ocand;         // [n_queries, ocap] candidates
  int32_t ocap;
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
// |approximate - fp32-exact| <= tc_filter_relerr(dim) * ||row|| * ||query||  for the filter scores.  Terms:
//   inputs   fp32 bank (kTcTf32x1): the tensor core truncates bank words to TF32 (relative error < 2^-10), queries are
//            rounded to nearest TF32 (<= 2^-11); by Cauchy-Schwarz the sum of the per-element errors is at most
//            (2^-10 + 2^-11 + 2^-21) * ||row|| * ||query||.  bf16 bank (kTcBf16x1): the bank is exact, queries are
//            rounded to nearest bf16 -- 8 significand bits, i.e. <= 2^-8 (NOT 2^-9).
//   2^-19    the product term above, and the fp32 roundings inside ||query|| and the cached ||row||^2 that E is built from.
//   dim*2^-22  everything that grows with the length of the sum: the tensor core adds dim/8 (dim/16) times into an fp32
//            accumulator with truncation -- modelled as <= 1 ulp per MMA (dim * 2^-26), but the hardware's internal
//            alignment of the 8 products is undocumented, so up to 9 ulp per MMA (2.25 * dim * 2^-24) are budgeted; the
//            fp32 re-scoring the bound is compared with rounds (dim/32 + 5) times to nearest, and so does each of the
//            two ||row||^2 sums (filter side / exact side) the normalised scores are divided by.
// The bound is deliberately not tight in its length term: E only widens the band of candidates refine.cu re-scores
// (a few rows per query), while a violated bound would silently lose a row.  refine.cu checks  u - 2E <= s <= u  for
// every candidate it re-scores (devqa_filter_bound_violations), tests/test_gpu_filter.py measures the actual error
// on constructed worst cases against this bound.
// fp16 SHADOW of an fp32 bank (shadow = true): rows and queries are both rounded to nearest fp16 -- 11 significand bits,
// <= 2^-11 each, after a per-row / per-query power-of-two scaling that keeps every value in fp16's normal range (see
// synth.cu shadow_convert_kernel; the few elements that still go subnormal add < 2^-31) -- so
// inputs = (1 + 2^-11)^2 - 1 = 2^-10 + 2^-22 relative to the ORIGINAL fp32 row's norm: tighter than the TF32 filter.
__host__ __device__ inline float tc_filter_relerr(int64_t dim, bool bf16_bank = false, bool shadow = false) {
  const float inputs = shadow ? 0.0009765625f + 2.384185791015625e-07f         // 2^-10 + 2^-22
                              : (bf16_bank ? 0.00390625f : 1.5f * 0.0009765625f);   // 2^-8 | 2^-10 + 2^-11
  return inputs + 1.9073486328125e-06f + float(dim) * 2.384185791015625e-07f;  // + 2^-19 + dim * 2^-22
}
// internal flag bit (never set by callers): the 2-byte bank being scanned is the scaled-fp16 shadow of an fp32 bank
constexpr unsigned kFlagShadowScan = 1u << 14;
bool scan_tc_supported(int bank_dtype, const void* bank, int64_t dim, int64_t bank_ld, int top_k);
int scan_tc_plan(int mode, int n_queries, int64_t dim, int kcap, bool cached_norms, int64_t n_tiles, int sm_count,
                 unsigned flags, int* qp_out, int* stages_out, bool* pair_out, bool oneshot = false, bool emit = false,
                 bool* resident_out = nullptr);
constexpr int kOneshotCap = 512;       // candidates per query the one-shot scan can emit (refine's first-cut capacity)
constexpr int kEmitCap = 128;          // candidates per (CTA, query) a long scan can emit (everything that ever beat the running
                                       // threshold: ~1500 per query over 148 CTAs with a bf16 shadow, far fewer with the fp16 one)
size_t scan_tc_qcat_bytes(int mode, int64_t dim, int q_pad);
bool scan_tc_exchanges_thresholds(int mode, int64_t dim);   // scan-and-queue epilogue + gmax exchange (else: seeded score-tile epilogue)
// queries -> normalised, split (hi/lo TF32; hi/mid/lo bf16; hi TF32 only for the filter), laid out per pass as the
// MMA B operand (see scan_tc.cu); also zeroes the shared threshold array gthr [q_pad].  Filter mode also writes,
// per query, the scale the normalisation applied (qinv) and the error scale qerr = tc_filter_relerr(dim) *
// ||query as used||, and zeroes seed_u [q_pad] and the threshold exchange table gmax [q_pad, gmax_stride].
int scan_tc_prep_queries(int mode, const float* queries, int64_t query_ld, int n_queries, int qp, int q_pad,
                         int64_t dim, unsigned flags, void* qcat, uint32_t* gthr, float* qinv, float* qerr,
                         uint32_t* seed_u, uint32_t* gmax, int gmax_stride, const uint32_t* need, cudaStream_t stream,
                         uint32_t* zero_words = nullptr, int n_zero = 0,   // filter modes: further words to zero
                         float* qscale = nullptr);                         // fp16 shadow: out [n_queries] 2^-f per query
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
  const uint32_t* cand_count; // optional [n_lists, n_queries]: valid entries per list (emit mode: lists are unsorted
                              // regions filled from the front); nullptr: sorted lists, empty slots are 0
  const uint32_t* gmax;       // [n_queries (padded), gmax_stride] each CTA's best upper bound per query (score codes)
  int32_t gmax_stride;
  const float* qerr;          // [n_queries] tc_filter_relerr(dim) * ||query as used||
  const uint32_t* bank_max_sumsq;   // un-normalised scores: max ||row||^2 (float bits)
  // refine_lists_launch, direct mode (one-shot scans): the candidates are already one unordered array per query
  const uint32_t* dcount;     // [n_queries] candidates emitted (may exceed dcap: then the query is flagged)
  const uint64_t* dcand;      // [n_queries, dcap]
  int32_t dcap;
  int32_t kp, top_k;
  uint32_t flags;
  IdMap idmap;
  const uint32_t* only_if;    // refine_launch only: when set, re-score (and re-sort) akeys of the flagged queries and
                              // nothing else -- no completeness verdict (used on the exact fallback's result)
  uint32_t* need;             // [n_queries] out: 1 when a lost row could still belong to the exact top-k
  uint32_t* fallbacks;        // device counter of such queries (devqa_filter_fallbacks)
  uint32_t* violations;       // device counter: re-scored candidates whose exact score lay outside [u - 2E, u]
  uint64_t* out_keys; float* out_scores; int64_t* out_ids;   // exactly one of keys / (scores, ids): [n_queries, top_k]
};
int refine_launch(const RefineParams& p, cudaStream_t stream);          // from the merged list akeys
int refine_lists_launch(const RefineParams& p, cudaStream_t stream);    // from the per-CTA lists (n_lists >= top_k)

// merge.cu: reduce `n_lists` sorted lists per query ([n_lists, n_queries, list_stride], the first
// list_len entries of each valid) to the final top_k.  Exactly one of (out_keys) or
// (out_scores, out_ids) is written: [n_queries, top_k].  With kth_codes != nullptr nothing else is written:
// kth_codes[q] receives the monotone score code of the k-th best key (0 when fewer than k exist).
// With need != nullptr only queries whose flag is set are touched.
//
// Peer-to-peer exchange fused into the merge (multi-GPU search): with `xchg` set, the CTA of query q first
// WRITES this rank's sorted top-k of q straight into every peer's mailbox over NVLink (plain stores into IPC-mapped
// memory, a system-scope fence, then a release-store of the call's sequence number into the peer's flag word), waits
// until all `world` flags of q in its own mailbox carry that sequence number, and then merges the `world` lists lying
// there -- ONE kernel instead of ncclAllGather + merge, no collective launch on the critical path.
constexpr int kP2pMaxWorld = 16;
constexpr int kP2pMaxQ = 512;                 // queries per call through the mailbox (all CTAs must be co-resident)
constexpr int64_t kP2pMaxKeys = 65536;        // keys per (slot, source rank): n_queries * top_k of a call
// The sequence number lives in DEVICE memory (header word 1 of the local mailbox; the kernel's last CTA advances it), so
// a search captured in a CUDA graph exchanges correctly on every replay; calls alternate between two slots.
struct P2pExchange {
  const uint64_t* local_keys;                 // [n_queries, list_stride] this rank's sorted top-k
  uint64_t* peer_data[kP2pMaxWorld];          // per rank: slot 0 of its data, a slot = [source rank][kP2pMaxKeys], used as [n_queries][list_stride]
  uint32_t* peer_flags[kP2pMaxWorld];         // per rank: slot 0 of its flags, a slot = [source rank][kP2pMaxQ]
  uint32_t* header;                           // local mailbox header: [0] timeouts, [1] exchanges completed, [2] arrival ticket
  int32_t rank, world;
};
int merge_lists_launch(const uint64_t* lists, int n_lists, int64_t n_queries, int list_stride,
                       int list_len, int top_k, uint64_t* out_keys, float* out_scores,
                       int64_t* out_ids, cudaStream_t stream, uint32_t* kth_codes = nullptr,
                       const uint32_t* need = nullptr, const P2pExchange* xchg = nullptr);

// select.cu: threshold-select (LiveEdit retrieve_moes): ordered compaction of the mask the stream kernel wrote.
struct SelectParams {
  const void* bank; int bank_dtype; int64_t n_rows, dim, bank_ld;
  const float* queries; int64_t query_ld; int32_t n_queries;
  unsigned flags; float scale; IdMap idmap;
  const uint8_t* mask; int64_t mask_ld;       // [n_queries, mask_ld]
  int64_t* out_ids; float* out_scores;        // [n_queries, cap]: selected rows in ascending row order (-1 / -inf beyond)
  int64_t cap;
  int64_t* out_counts;                        // [n_queries] number of selected rows (may exceed cap)
};
int select_compact_launch(const SelectParams& p, cudaStream_t stream);

// gate.cu: HICE scope gate -- logits = relu(q/||q|| @ W_rand) @ Wo^T, argmax, m-loc memory override.
struct GateParams {
  const float* queries; int64_t query_ld; int32_t n_queries; int32_t dim;
  const float* w_rand; int64_t w_ld; int32_t m;        // [dim, m]
  const float* wo; int64_t wo_ld; int32_t n_classes;   // [n_classes, m]
  unsigned flags;                                      // DEVQA_NORMALIZE_QUERY / DEVQA_NOEPS
  float* partial;                                      // scratch [scope_gate_partial_floats()]
  uint32_t* ticket;                                    // zeroed device word (re-zeroed by the kernel)
  float* logits; int32_t* preds;                       // out [n_queries, n_classes], [n_queries]
  const float* mem_scores; int64_t mem_stride; float mem_threshold;   // optional: best memory similarity per probe
};
size_t scope_gate_partial_floats(const GateParams& p);
int scope_gate_launch(const GateParams& p, cudaStream_t stream);

// vmm.cu: bank storage that grows in place (virtual memory management); nullptr / DEVQA_E_UNSUPPORTED when unavailable
struct VmmBuffer;
VmmBuffer* vmm_create(int device);
int vmm_grow(VmmBuffer* b, size_t want_bytes, bool* moved);
void* vmm_data(const VmmBuffer* b);
size_t vmm_mapped(const VmmBuffer* b);
void vmm_destroy(VmmBuffer* b);

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
// merge.cu: one 1024-wide pass of a deep top-k: [n_queries, kj] keys -> columns [col0, col0 + kj) of the caller's
// [n_queries, top_k] outputs (keys, or scores + ids), and below[q] = the pass's last key (1 when the bank is exhausted).
int deep_emit_launch(const uint64_t* pass_keys, int64_t n_queries, int kj, int top_k, int col0, uint64_t* out_keys,
                     float* out_scores, int64_t* out_ids, uint64_t* below, cudaStream_t stream);
// synth.cu helper kernels
int widen_rows_launch(const void* src, int64_t src_ld, float* dst, int64_t dst_ld, int64_t n_rows, int64_t dim,
                      cudaStream_t stream);
int convert_rows_launch(const float* src, int64_t src_ld, void* dst, int dst_dtype, int64_t dst_ld,
                        int64_t n_rows, int64_t dim, cudaStream_t stream);
// fp32 rows -> per-row scaled fp16 shadow rows + the scale to undo it (synth.cu)
int shadow_convert_launch(const float* src, int64_t src_ld, void* dst, int64_t dst_ld, float* scale_out, int64_t n_rows,
                          int64_t dim, cudaStream_t stream);

}  // namespace devqa
