/*
 * devqa.h -- C ABI of the B200-native fact-bank retrieval library (libdevqa.so).
 *
 * This is the drop-in boundary for ONE hot path of sev777/DE-VQA ("LiveEdit"):
 * the fact-bank similarity search  L2-normalise -> query x bank^T -> row-wise top-k
 * that the reference reaches through sentence_transformers.util.  Citations are
 * relative to /root/reference/DE-VQA.
 *
 *   reference interface                                   replaced by
 *   ----------------------------------------------------  --------------------------------
 *   util.normalize_embeddings(bank)      dataset/vllm.py:104,117; ike_main.py:37,194
 *   bank / torch.norm(bank)  (no eps)    ike_main.py:238; multimodal_evaluate.py:410;
 *                                        HICE.py:170            -> flags of devqa_*search*
 *   util.dot_score(q, bank)              passed at every search site (torch.mm)
 *   util.semantic_search(q, bank, dot_score, top_k)
 *                                        dataset/vllm.py:68; ike_main.py:44,202,240;
 *                                        multimodal_evaluate.py:412; HICE.py:177;
 *                                        easyeditor/dataset/vqa.py:176; coco_caption.py:174
 *                                                               -> devqa_search_raw /
 *                                                                  devqa_bank_search*
 *   torch.tensor(pkl['embeddings']).to(dev)   dataset/vllm.py:102,116; ike_main.py:36,193
 *                                                               -> devqa_bank_create/append
 *   torch.cat([pool, new], 0) per edit   lte_vl.py:150; serac_vl.py:104; recipe_vl.py:163
 *                                                               -> devqa_bank_append (in place)
 *
 * Conventions: plain C types only; every function returns 0 (DEVQA_OK) or a
 * negative DEVQA_E_* code and leaves a message for devqa_last_error() (thread
 * local).  Unless a name ends in _host, pointer arguments are DEVICE pointers on
 * the bank's device, and the call only enqueues work on `stream` (a cudaStream_t
 * passed as void*; NULL = the legacy default stream) without synchronising.
 * There is no CPU fallback anywhere in the library.
 *
 * Result contract (all search entry points): for each query the min(top_k, N)
 * best rows, score-descending; ties broken by ascending row id (the reference's
 * tie order is implementation-defined).  Scores are fp32.
 * Output arrays are [n_queries, top_k] row-major; when N < top_k the trailing
 * slots of each row hold score = -inf and id = -1.
 */
#ifndef DEVQA_H_
#define DEVQA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DEVQA_VERSION 100  /* 0.1.0 */

#if defined(__GNUC__)
#define DEVQA_API __attribute__((visibility("default")))
#else
#define DEVQA_API
#endif

/* error codes */
#define DEVQA_OK 0
#define DEVQA_E_INVALID (-1)     /* bad argument */
#define DEVQA_E_CUDA (-2)        /* a CUDA runtime call failed */
#define DEVQA_E_NOMEM (-3)       /* device/host allocation failed */
#define DEVQA_E_UNSUPPORTED (-4) /* e.g. top_k > DEVQA_MAX_TOPK */
#define DEVQA_E_NCCL (-5)        /* NCCL missing or a collective failed */
#define DEVQA_E_STATE (-6)       /* object used in the wrong state */

#define DEVQA_MAX_TOPK 1024 /* results per pass: single-GPU searches accept any top_k (deeper ranks come 1024 at a time, one
                               bank pass each); devqa_topk_merge and the sharded search are limited to this */

/* bank element types */
#define DEVQA_F32 0
#define DEVQA_BF16 1 /* bank rows stored as bf16 (converted on append); scores still fp32 */

/* search flags */
#define DEVQA_NORMALIZE_BANK (1u << 0)  /* score /= max(||row||_2, 1e-12)  (fused; bank is read once) */
#define DEVQA_NORMALIZE_QUERY (1u << 1) /* query /= max(||q||_2, 1e-12)                                 */
#define DEVQA_NOEPS (1u << 2)           /* use x/||x|| with no clamp (HICE call sites); zero row -> NaN,
                                           and NaN ranks above every number, as torch.topk does        */
#define DEVQA_FORCE_STREAM (1u << 8)    /* testing/benchmark: force the bank-streaming (CUDA-core) kernel */
#define DEVQA_FORCE_TENSOR (1u << 9)    /* testing/benchmark: force the tcgen05 contraction kernel       */
#define DEVQA_FORCE_TF32X3 (1u << 10)   /* testing/benchmark: tcgen05 path without the TF32 filter + exact
                                           re-scoring stage (every score straight from the 3xTF32 split)  */

#define DEVQA_NO_SHADOW (1u << 15)      /* search an fp32 bank that has a shadow as if it had none */
#define DEVQA_FORCE_FILTER (1u << 11)   /* testing/benchmark: TF32 filter + exact re-scoring also on banks the
                                           library would scan with 3xTF32 alone (fewer than 8 tiles per SM)     */
#define DEVQA_DEBUG_TIGHT_BOUND (1u << 12) /* tests: run the filter with 1/64 of its error bound, so that the bound IS
                                              violated and the canary (devqa_filter_bound_violations) has to catch it */
#define DEVQA_DEBUG_NO_ONESHOT (1u << 13)  /* tests/benchmarks: short scans through the persistent kernel's list epilogue
                                              instead of the one-shot mode (one tile per CTA, grid-wide threshold exchange) */
#define DEVQA_DEBUG_NO_EMIT (1u << 3)      /* tests/benchmarks: long filter scans keep per-CTA sorted lists (the round-1 epilogue)
                                              instead of appending survivors to the flat candidate array */
#define DEVQA_DEBUG_FILTER_ONLY (1u << 22) /* tests: return the TF32 filter's approximate top-k as is (no exact
                                              re-scoring) so that its error bound can be checked            */

/* Environment knob (read once): DEVQA_FUSED_MERGE=0 makes the small-batch stream kernel launch a separate merge
 * kernel instead of letting its last CTA merge the per-CTA lists (the default since it saves a launch per search). */

typedef struct devqa_bank devqa_bank; /* opaque, library-owned fact bank */
typedef struct devqa_comm devqa_comm; /* opaque, one NCCL communicator + staging */

/* ---- library ------------------------------------------------------------------------- */
DEVQA_API int devqa_version(void);
DEVQA_API const char* devqa_last_error(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
DEVQA_API int64_t devqa_launch_count(void);

/* Per-kernel timing for benchmarks: while enabled, every search records a CUDA-event pair on its
 * own stream around its scan launches (the dominant kernel); collect() synchronises, returns the
 * summed scan time, the number of scan launches and searches since the last collect, and resets. */
DEVQA_API int devqa_profile_enable(int device, int enable);
DEVQA_API int devqa_profile_collect(int device, double* scan_ms_total, int64_t* scan_launches,
                                    int64_t* searches);

/* fp32 banks, batches > 4 queries: the tensor-core scan filters with ONE TF32 product, re-scores its
 * candidates exactly in fp32 and proves per query that the exact top-k is among them; queries for which the
 * proof fails (heavily duplicated / clustered rows, NaN) are redone by the fp32-exact 3xTF32 scan in the same
 * call.  Returns (and resets) how many queries took that fallback on `device`; synchronises the device. */
DEVQA_API int devqa_filter_fallbacks(int device, int64_t* count);
/* The filter's proof rests on  u - 2E <= s <= u  (u: the scan's upper bound of a row's score, s: its exact fp32 score,
 * E = devqa_filter_relerr(dim) * ||row|| * ||query||).  The re-scoring kernels check it for every candidate they
 * re-score; a violation sends the query to the exact fallback and is counted here.  Returns (and resets) that counter;
 * synchronises the device.  Expected value: 0, always. */
DEVQA_API int devqa_filter_bound_violations(int device, int64_t* count);
/* The relative error bound the filter uses for `dim` columns on an fp32 (DEVQA_F32) or bf16 (DEVQA_BF16) bank. */
DEVQA_API double devqa_filter_relerr(int64_t dim, int bank_dtype);

/* Debugging aid: when set to a device buffer of 16 x 256 int64, CTA 0 of the tcgen05 scan records
 * clock64() samples of its warp roles (producer, split, MMA, epilogue) into it.  NULL disables. */
DEVQA_API int devqa_debug_tc_trace(int64_t* device_buf_16x256);

/* ---- stateless search over a caller-owned bank -----------------------------------------
 * The literal replacement of  semantic_search(q, corpus, score_function=dot_score, top_k)
 * on CUDA tensors: `bank` is [n_rows, dim] with row stride `bank_ld` elements (>= dim),
 * `queries` is [n_queries, dim] fp32 with row stride `query_ld`.  Returned ids are
 * id_base + row.  */
DEVQA_API int devqa_search_raw(const void* bank, int bank_dtype, int64_t n_rows, int64_t dim, int64_t bank_ld,
                     const float* queries, int64_t n_queries, int64_t query_ld, int top_k,
                     unsigned flags, int64_t id_base, float* out_scores, int64_t* out_ids,
                     void* stream);

/* The same with side information about the caller-owned bank: row_sumsq [n_rows] = ||row||^2 per row and
 * max_sumsq_bits = the bit pattern of their maximum (both DEVICE memory, e.g. from devqa_row_sumsq, NULL = none).
 * With it an fp32 bank is searched exactly like a library-owned one -- the TF32 filter + exact re-scoring also for
 * un-normalised dot scores (the reference's  normalise first, then semantic_search(q, bank, dot_score)  pattern,
 * dataset/vllm.py:66-68,104).  Results are bit-identical to devqa_search_raw's; the side information only decides
 * how fast they arrive, but it MUST describe the bank as it is now (a stale array can lose rows). */
DEVQA_API int devqa_search_raw_ex(const void* bank, int bank_dtype, int64_t n_rows, int64_t dim, int64_t bank_ld,
                        const float* queries, int64_t n_queries, int64_t query_ld, int top_k,
                        unsigned flags, int64_t id_base, const float* row_sumsq,
                        const uint32_t* max_sumsq_bits, float* out_scores, int64_t* out_ids, void* stream);
/* ... and with a 2-byte SHADOW of the caller-owned fp32 rows (devqa_make_shadow; both NULL = none): large scans with
 * top_k <= 64 stream the shadow (half the bytes) and re-score their candidates against the fp32 rows -- results stay
 * bit-identical to devqa_search_raw's.  The shadow needs row_sumsq / max_sumsq_bits as well and, like them, MUST describe
 * the bank as it is now.  This is what devqa_bank_set_shadow does for a library-owned bank; Level-0 callers
 * (de-vqa_b200/util.py: the reference's semantic_search(q, corpus) over the corpus tensor it already holds,
 * dataset/vllm.py:68, ike_main.py:202) cache it per corpus tensor. */
DEVQA_API int devqa_search_raw_shadowed(const void* bank, int bank_dtype, int64_t n_rows, int64_t dim, int64_t bank_ld,
                              const float* queries, int64_t n_queries, int64_t query_ld, int top_k,
                              unsigned flags, int64_t id_base, const float* row_sumsq,
                              const uint32_t* max_sumsq_bits, const void* shadow, int64_t shadow_ld,
                              const float* shadow_scale, float* out_scores, int64_t* out_ids, void* stream);
/* Write the shadow of fp32 rows: shadow [n_rows, shadow_ld] 2-byte elements (scaled fp16: each row multiplied by a power
 * of two so that its largest magnitude lands in [2^13, 2^14)), shadow_scale [n_rows] the inverse factors.  shadow_ld >=
 * dim and a multiple of 8, shadow 16-byte aligned.  All DEVICE memory. */
DEVQA_API int devqa_make_shadow(const float* bank, int64_t n_rows, int64_t dim, int64_t bank_ld, void* shadow,
                                int64_t shadow_ld, float* shadow_scale, void* stream);
/* ||row||^2 of every row of a caller-owned bank (one warp per row): out_sumsq [n_rows]; out_max_bits (optional) is a
 * device word the caller zeroed, which receives the bit pattern of the largest sum (atomicMax). */
DEVQA_API int devqa_row_sumsq(const void* bank, int bank_dtype, int64_t n_rows, int64_t dim, int64_t bank_ld,
                              float* out_sumsq, uint32_t* out_max_bits, void* stream);

/* Stand-alone row normalisation, dst = src / max(||src||_2, 1e-12) (or / ||src|| with
 * DEVQA_NOEPS): util.normalize_embeddings (dataset/vllm.py:66,104,117).  dst may equal src.
 * The search entry points fuse normalisation and do not need this.                        */
DEVQA_API int devqa_normalize_rows(const float* src, int64_t src_ld, float* dst, int64_t dst_ld, int64_t n_rows,
                         int64_t dim, unsigned flags, void* stream);

/* ---- library-owned, appendable bank ------------------------------------------------------
 * Replaces the per-request pickle -> numpy -> torch.tensor -> .to(device) upload and the
 * O(N) torch.cat append of the lifelong editors with a resident, in-place growable buffer.
 * capacity_rows is a hint (the bank grows geometrically in place when exceeded).          */
DEVQA_API int devqa_bank_create(devqa_bank** out, int device, int64_t dim, int64_t capacity_rows, int dtype);
DEVQA_API int devqa_bank_destroy(devqa_bank* bank);
DEVQA_API int64_t devqa_bank_size(const devqa_bank* bank);     /* rows currently stored, <0 on error */
DEVQA_API int64_t devqa_bank_capacity(const devqa_bank* bank); /* rows that fit without growing      */
DEVQA_API int64_t devqa_bank_dim(const devqa_bank* bank);
DEVQA_API int devqa_bank_dtype(const devqa_bank* bank);
DEVQA_API int devqa_bank_device(const devqa_bank* bank);
DEVQA_API const void* devqa_bank_data(const devqa_bank* bank); /* device pointer to row 0 */
DEVQA_API int devqa_bank_reserve(devqa_bank* bank, int64_t capacity_rows, void* stream);
/* Append n fp32 rows (row stride `ld` elements).  rows_is_host != 0: `rows` is host memory
 * and the call stages + copies it (synchronous w.r.t. the host buffer).                   */
DEVQA_API int devqa_bank_append(devqa_bank* bank, const float* rows, int64_t n, int64_t ld, int rows_is_host,
                      void* stream);
/* Optional 2-byte SHADOW of an fp32 bank: fp16 rows, each scaled by its own power of two (kept per row) so that any
 * finite fp32 row fits fp16's normal range.  The bank stays fp32 and every result stays the exact fp32 result
 * (bit-identical to a bank without a shadow); but scans stream the 2-byte shadow instead of the 4-byte rows: the
 * tensor-core filter runs on it with an error bound that also covers the rounding of the rows (2^-10 in all, tighter
 * than the TF32 filter's 1.5 * 2^-10), its candidates are re-scored against the fp32 rows, and the usual proof /
 * canary / exact fallback applies.  Costs 50% more device memory and one extra conversion pass per append; halves the
 * bytes of every HBM-bound search and doubles the tensor rate of the large-batch ones.  enable = 0 drops it. */
DEVQA_API int devqa_bank_set_shadow(devqa_bank* bank, int enable, void* stream);
DEVQA_API int devqa_bank_has_shadow(const devqa_bank* bank);
/* Copy rows [first_row, first_row + n) out as fp32 into a DEVICE buffer (bf16 banks are widened). */
DEVQA_API int devqa_bank_read_rows(const devqa_bank* bank, int64_t first_row, int64_t n, float* dst,
                                   int64_t dst_ld, void* stream);
/* Drop rows from the end (restore_to_original_model of the lifelong editors). */
DEVQA_API int devqa_bank_truncate(devqa_bank* bank, int64_t n_rows);
/* How local rows map to global corpus ids (multi-GPU sharding):
 *   global = id_base + ((row / chunk_rows) * world + rank) * chunk_rows + row % chunk_rows
 * Contiguous shard: chunk_rows = 0 (=> global = id_base + row).                           */
DEVQA_API int devqa_bank_set_id_map(devqa_bank* bank, int64_t id_base, int64_t chunk_rows, int rank, int world);

DEVQA_API int devqa_bank_search(devqa_bank* bank, const float* queries, int64_t n_queries, int64_t query_ld,
                      int top_k, unsigned flags, float* out_scores, int64_t* out_ids, void* stream);
/* Same, with HOST queries and HOST outputs: stages through pinned memory, copies in, runs,
 * copies out and synchronises -- the one D2H + sync the reference's .cpu().tolist() also pays. */
DEVQA_API int devqa_bank_search_host(devqa_bank* bank, const float* queries_host, int64_t n_queries,
                           int64_t query_ld, int top_k, unsigned flags, float* out_scores_host,
                           int64_t* out_ids_host, void* stream);
/* Local top-k as order-preserving packed keys, [n_queries, top_k] uint64:
 *   key = (monotone(score bits) << 32) | (0xFFFFFFFF - global_id); empty slot = 0.        */
DEVQA_API int devqa_bank_search_packed(devqa_bank* bank, const float* queries, int64_t n_queries,
                             int64_t query_ld, int top_k, unsigned flags, uint64_t* out_keys,
                             void* stream);

/* Threshold-select instead of top-k (LiveEdit's hard routing, editor/vllm_editors/liveedit/liveedit.py:132-143:
 * retrieval_map = vis_sim > vis_sim_prot over the whole pool): every row whose  score * scale > thresholds[q].
 * One pass of the bank-streaming kernel writes the boolean mask (out_mask [n_queries, size] bytes, optional), a second
 * small kernel compacts it: out_ids / out_scores [n_queries, cap] receive the selected rows in ASCENDING id order with
 * their scaled scores (-1 / -inf beyond the count; rows beyond `cap` are only in the mask), out_counts [n_queries] the
 * number of selected rows (may exceed cap).  `thresholds` is a DEVICE array: no host synchronisation, no k cap.
 * NaN scores are not selected (torch's > is false for NaN).  Scores are bit-identical to devqa_bank_search's. */
DEVQA_API int devqa_bank_select_above(devqa_bank* bank, const float* queries, int64_t n_queries, int64_t query_ld,
                                      const float* thresholds, float scale, unsigned flags, uint8_t* out_mask,
                                      int64_t* out_ids, float* out_scores, int64_t cap, int64_t* out_counts,
                                      void* stream);

/* ---- HICE scope gate (easyeditor/evaluate/multimodal_evaluate.py:382-392; easyeditor/trainer/algs/HICE.py:238-263) ----
 * logits = relu(q/||q|| @ W_rand) @ Wo^T, preds = argmax(logits); W_rand [dim, m] (row stride w_ld), Wo [n_classes, m]
 * (row stride wo_ld), all DEVICE fp32.  `flags`: DEVQA_NORMALIZE_QUERY | DEVQA_NOEPS reproduces the reference's
 * q / torch.norm(q).  mem_scores (optional, device [n_queries]): best cosine similarity of each probe against the m-loc
 * memory; a probe with mem_scores > mem_threshold is out of scope (preds = 0), as in run_classifier2.
 * One kernel for all probes; deterministic (fixed-order reductions).  Works on the current device. */
DEVQA_API int devqa_scope_gate(const float* queries, int64_t n_queries, int64_t query_ld, int64_t dim,
                               const float* w_rand, int64_t w_ld, int64_t m, const float* wo, int64_t wo_ld,
                               int n_classes, unsigned flags, const float* mem_scores, float mem_threshold,
                               float* out_logits, int32_t* out_preds, void* stream);
/* The whole per-probe retrieval of compute_hice_multimodal_edit_quality (multimodal_evaluate.py:503-573) for all
 * probes of a request in ONE call on one stream: (1) the m-loc memory check -- no-eps cosine top-1 of mem_queries
 * (NULL: of `queries`) against `memory` (NULL or empty: skipped); (2) the RanPAC scope gate on `queries` with the
 * memory override; (3) the in-context lookup -- top_k search of `queries` over `bank` with `flags` (the reference uses
 * DEVQA_NORMALIZE_BANK | DEVQA_NORMALIZE_QUERY | DEVQA_NOEPS, select_icl_samples :404-412).  out_mem_sim (optional)
 * [n_queries]; out_logits [n_queries, n_classes]; out_preds [n_queries]; out_scores / out_ids [n_queries, top_k]
 * (returned for every probe; callers use the in-scope ones, preds != 0).  No host synchronisation. */
DEVQA_API int devqa_bank_hice_probe(devqa_bank* bank, devqa_bank* memory, const float* queries,
                                    const float* mem_queries, int64_t n_queries, int64_t query_ld,
                                    const float* w_rand, int64_t w_ld, int64_t m, const float* wo, int64_t wo_ld,
                                    int n_classes, float mem_threshold, int top_k, unsigned flags,
                                    float* out_logits, int32_t* out_preds, float* out_mem_sim,
                                    float* out_scores, int64_t* out_ids, void* stream);

/* ---- multi-GPU merge ---------------------------------------------------------------------
 * gathered: [n_lists, n_queries, top_k] packed keys, each [top_k] run sorted descending. */
DEVQA_API int devqa_topk_merge(const uint64_t* gathered, int n_lists, int64_t n_queries, int top_k,
                     float* out_scores, int64_t* out_ids, void* stream);

/* One NCCL communicator per process (one process per GPU).  Rank 0 calls
 * devqa_dist_unique_id, ships the 128 bytes to every rank by any host channel
 * (torch.distributed, MPI, a file), then every rank calls devqa_dist_init.               */
DEVQA_API int devqa_dist_unique_id(void* out_id_128_bytes);
DEVQA_API int devqa_dist_init(devqa_comm** out, const void* id_128_bytes, int rank, int world, int device);
DEVQA_API int devqa_dist_destroy(devqa_comm* comm);
/* Peer-to-peer mailboxes (optional, one box): every rank exports a 64-byte IPC handle of its mailbox, the handles travel
 * over the same host channel as the NCCL id, and every rank connects with the `world` handles in rank order.  After
 * that, searches of up to 512 queries (n_queries * top_k <= 65536) exchange their local top-k by writing straight into
 * the peers' mailboxes over NVLink from inside the merge kernel -- scan -> re-scoring -> ONE push/wait/merge kernel, no
 * collective launch.  Larger batches keep the ncclAllGather.  connect() returns DEVQA_E_UNSUPPORTED when a peer's
 * memory cannot be mapped (the communicator then simply stays on NCCL); all ranks must end up with the same setting
 * (devqa_dist_p2p_set).  devqa_dist_p2p_timeouts: waits that gave up after ~10 s (a peer died); expected 0. */
DEVQA_API int devqa_dist_p2p_export(devqa_comm* comm, void* out_handle_64_bytes);
DEVQA_API int devqa_dist_p2p_connect(devqa_comm* comm, const void* handles_world_x_64_bytes);
DEVQA_API int devqa_dist_p2p_set(devqa_comm* comm, int enable);
DEVQA_API int devqa_dist_p2p_enabled(const devqa_comm* comm);
DEVQA_API int devqa_dist_p2p_timeouts(devqa_comm* comm, int64_t* count);
DEVQA_API int devqa_dist_rank(const devqa_comm* comm);
DEVQA_API int devqa_dist_world(const devqa_comm* comm);
/* scan of the local shard -> ONE ncclAllGather of [n_queries, top_k] packed keys ->
 * fused merge; every rank receives the full result.  Queries are replicated.             */
DEVQA_API int devqa_bank_search_dist(devqa_bank* bank, devqa_comm* comm, const float* queries,
                           int64_t n_queries, int64_t query_ld, int top_k, unsigned flags,
                           float* out_scores, int64_t* out_ids, void* stream);
DEVQA_API int devqa_bank_search_dist_host(devqa_bank* bank, devqa_comm* comm, const float* queries_host,
                                int64_t n_queries, int64_t query_ld, int top_k, unsigned flags,
                                float* out_scores_host, int64_t* out_ids_host, void* stream);

/* ---- synthetic data (benchmarks, tests) ---------------------------------------------------
 * Counter-based generator keyed by (seed, global row, column) so any sharding reproduces
 * the same bank.  kind 0: unit-norm N(0,1) rows (MiniLM-like); kind 1: un-normalised
 * "hidden-state" rows, N(0,1) * exp(N(0,0.5^2)) per row with 4 outlier columns x50.       */
DEVQA_API int devqa_fill_synthetic(float* dst, int64_t n_rows, int64_t dim, int64_t ld, uint64_t seed,
                         int64_t first_global_row, int kind, void* stream);
DEVQA_API int devqa_bank_append_synthetic(devqa_bank* bank, int64_t n_rows, uint64_t seed,
                                int64_t first_global_row, int kind, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DEVQA_H_ */
