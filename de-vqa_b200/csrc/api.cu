// C ABI of libdevqa (include/devqa.h): bank object, workspace, kernel selection, NCCL plumbing.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <mutex>
#include <new>

#include "scan.cuh"
#include "topk.cuh"
#include "nccl_dyn.h"

namespace devqa {

// ---- errors / counters --------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
bool pdl_enabled() {
  static const bool on = getenv("DEVQA_PDL") == nullptr || atoi(getenv("DEVQA_PDL")) != 0;
  return on;
}

// ---- per-device scratch ---------------------------------------------------------------------------
// Candidate lists of the scan kernels, pinned staging for the _host entry points.  One search runs
// per device at a time from one host thread (the reference's usage); work on a
// different stream than the previous call is ordered behind it with an event.
struct Scratch {
  std::mutex mu;
  int sm_count = 0;
  uint64_t* cand = nullptr;   size_t cand_bytes = 0;
  void* dev_io = nullptr;     size_t dev_io_bytes = 0;     // device queries / outputs of _host calls
  void* pin_io = nullptr;     size_t pin_io_bytes = 0;     // pinned mirror
  uint64_t* keys = nullptr;   size_t keys_bytes = 0;       // packed local top-k (dist)
  uint64_t* gathered = nullptr; size_t gathered_bytes = 0; // all-gather target (dist)
  void* tiny = nullptr;       size_t tiny_bytes = 0;       // padded copy of banks shorter than one row group
  uint8_t* selmask = nullptr; size_t selmask_bytes = 0;    // threshold-select: [Q, N] mask when the caller wants none
  uint64_t* deep = nullptr;   size_t deep_bytes = 0;       // deep top-k: one pass's keys [Q, 1024] + per-query bound keys
  float* qsplit = nullptr;    size_t qsplit_bytes = 0;     // TF32 hi / lo query planes of the tcgen05 scan
  uint64_t* akeys = nullptr;  size_t akeys_bytes = 0;      // filter scan: merged approximate top-K' lists + per-query side arrays
  uint32_t* fallbacks = nullptr;                           // device counter: queries the exact fallback had to redo
  uint32_t* ticket = nullptr;                              // device word: arrival counter of the stream kernel's fused merge
  uint32_t* violations = nullptr;                          // device counter: re-scored candidates outside [u - 2E, u] (refine.cu canary)
  uint32_t* gate_ticket = nullptr;                         // device word: arrival counter of the scope-gate kernel
  float* gate = nullptr;      size_t gate_bytes = 0;       // scope gate: per-CTA partial logits + memory-search results
  cudaStream_t last_stream = nullptr;
  cudaEvent_t done = nullptr;
  // optional per-kernel timing (devqa_profile_*): event pairs around the scan passes of each search
  bool profiling = false;
  static constexpr int kProfSlots = 8192;
  cudaEvent_t* prof_ev = nullptr;     // [2 * kProfSlots]
  int prof_used = 0;
  int64_t prof_scan_launches = 0;
  int64_t prof_searches = 0;
};
static Scratch g_scratch[64];

struct DeviceGuardLite {
  int prev = -1;
  explicit DeviceGuardLite(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); }
  ~DeviceGuardLite() { if (prev >= 0) cudaSetDevice(prev); }
};

static thread_local cudaStream_t tl_stream = nullptr;   // the stream of the entry point this thread is inside (scratch_enter)

static int grow(void** p, size_t* have, size_t want, bool pinned) {
  if (*have >= want) return DEVQA_OK;
  {
    // (re)allocation synchronises the device, which is illegal while the call's stream is being captured: say so
    // instead of failing inside cudaDeviceSynchronize and invalidating the capture.  A search is capturable once an
    // identical (or larger) call has run eagerly and sized the scratch -- the warm-up graph capture needs anyway.
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(tl_stream, &st);
    if (st != cudaStreamCaptureStatusNone) {
      set_error("scratch of %zu bytes must grow during stream capture: run the same call once eagerly before capturing", want);
      return DEVQA_E_STATE;
    }
  }
  if (*p) {
    DEVQA_CUDA(cudaDeviceSynchronize());
    if (pinned) cudaFreeHost(*p); else cudaFree(*p);
    *p = nullptr; *have = 0;
  }
  size_t sz = want + want / 4 + 4096;
  cudaError_t e = pinned ? cudaMallocHost(p, sz) : cudaMalloc(p, sz);
  if (e != cudaSuccess) {
    set_error("%s of %zu bytes failed: %s", pinned ? "cudaMallocHost" : "cudaMalloc", sz, cudaGetErrorString(e));
    return DEVQA_E_NOMEM;
  }
  *have = sz;
  return DEVQA_OK;
}

static std::mutex g_scratch_init_mu;

static int scratch_for(int device, Scratch** out) {
  DEVQA_REQUIRE(device >= 0 && device < 64, "device %d out of range", device);
  Scratch* s = &g_scratch[device];
  std::lock_guard<std::mutex> init_lock(g_scratch_init_mu);   // lazy init from any host thread
  if (s->sm_count == 0) {
    int n = 0;
    DEVQA_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device));
    int major = 0;
    DEVQA_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    if (major != 10) {
      set_error("libdevqa is built for sm_100a only; device %d has compute capability %d.x", device, major);
      return DEVQA_E_UNSUPPORTED;
    }
    {
      // everything below belongs to `device`, whichever device the calling thread has current (an event created on
      // another device cannot be recorded on this device's streams)
      DeviceGuardLite guard(device);
      DEVQA_CUDA(cudaEventCreateWithFlags(&s->done, cudaEventDisableTiming));
      DEVQA_CUDA(cudaMalloc(&s->fallbacks, 16));
      DEVQA_CUDA(cudaMemset(s->fallbacks, 0, 16));
      s->ticket = s->fallbacks + 1;
      s->violations = s->fallbacks + 2;
      s->gate_ticket = s->fallbacks + 3;
    }
    s->sm_count = n;
  }
  *out = s;
  return DEVQA_OK;
}

// order this call behind the previous user of the scratch when the stream changed
static int scratch_enter(Scratch* s, cudaStream_t stream) {
  tl_stream = stream;
  if (s->last_stream != stream && s->done) {
    // (a capturing stream must not wait on an event from outside its capture; whoever captures a search has
    // synchronised with earlier work already, as CUDA graph capture requires)
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(stream, &st);
    if (st == cudaStreamCaptureStatusNone) DEVQA_CUDA(cudaStreamWaitEvent(stream, s->done, 0));
  }
  return DEVQA_OK;
}
static int scratch_leave(Scratch* s, cudaStream_t stream) {
  s->last_stream = stream;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(stream, &st);
  if (st == cudaStreamCaptureStatusNone) DEVQA_CUDA(cudaEventRecord(s->done, stream));
  return DEVQA_OK;
}

// ---- the search pipeline: scan passes -> cross-CTA merge ---------------------------------------------
struct SearchArgs {
  const void* bank; int dtype; int64_t n_rows, dim, bank_ld;
  const float* queries; int64_t n_queries, query_ld;
  int top_k; unsigned flags; IdMap idmap;
  const float* row_sumsq = nullptr;
  const uint32_t* max_sumsq = nullptr;   // library-owned banks: max ||row||^2 over every row ever appended (float bits)
  const uint64_t* below = nullptr;       // deep top-k passes: per-query exclusive upper bound key
  const void* shadow = nullptr;          // fp32 library banks: optional scaled-fp16 copy of the rows (the filter scans IT)
  int64_t shadow_ld = 0;
  const float* shadow_scale = nullptr;   // its per-row 2^-e
  const float* row_scale = nullptr;      // (set on the args of the filter scan over the shadow)
};

static int validate(const SearchArgs& a) {
  DEVQA_REQUIRE(a.dim > 0 && a.dim <= (1 << 20), "dim %lld unsupported", (long long)a.dim);
  DEVQA_REQUIRE(a.n_rows >= 0 && a.n_queries >= 0, "negative sizes");
  DEVQA_REQUIRE(a.bank_ld >= a.dim && a.query_ld >= a.dim, "row stride smaller than dim");
  DEVQA_REQUIRE(a.top_k >= 1, "top_k must be >= 1");
  DEVQA_REQUIRE(a.dtype == DEVQA_F32 || a.dtype == DEVQA_BF16, "unknown bank dtype %d", a.dtype);
  DEVQA_REQUIRE(a.top_k <= (1 << 24), "top_k %d unsupported", a.top_k);
  DEVQA_REQUIRE(a.n_rows == 0 || a.bank != nullptr, "bank pointer is null");
  DEVQA_REQUIRE(a.n_queries == 0 || a.queries != nullptr, "queries pointer is null");
  DEVQA_REQUIRE(a.idmap.base + a.n_rows * (a.idmap.chunk > 0 ? a.idmap.world : 1) <= 0xffffffffll,
                "corpus ids beyond 2^32 are not supported");
  return DEVQA_OK;
}

constexpr int kGmaxStride = 160;   // >= SM count, 16-byte multiple

// Side arrays of the filter scan (device), carved out of Scratch::akeys.
struct FilterSide {
  uint64_t* akeys = nullptr;     // [Q, K'] merged candidates (keys on the upper bound u)
  uint32_t* need = nullptr;      // [Q] set by refine: the exact fallback has to redo this query
  float* qinv = nullptr;         // [Q] scale applied by the query normalisation
  float* qerr = nullptr;         // [Q] tc_filter_relerr(dim) * ||query as used||
  float* qscale = nullptr;       // [Q] fp16 shadow: 2^-f of each query's scaling
  uint32_t* seed_u = nullptr;    // [Q padded] k-th best upper bound of the prefix sample (score codes)
  uint32_t* lostu = nullptr;     // [grid, Q] per-CTA bound on the rows its full list lost
  uint32_t* gmax = nullptr;      // [Q padded, kGmaxStride] per-CTA best upper bound per query (threshold exchange)
  int kp = 0;                    // K': capacity of the merged candidate list
  bool merge = true;             // in: merge the per-CTA lists into akeys (false: refine reads the lists themselves)
  // one-shot mode (short scans, one tile per CTA): the scan emits its candidates straight into ocand
  uint32_t* gbar = nullptr;      // [2 * 64] arrival counters  } zeroed together by the query-prep kernel
  uint32_t* ocount = nullptr;    // [Q] candidates per query   }
  uint64_t* ocand = nullptr;     // [Q, kOneshotCap]
  bool oneshot = false;          // out: the scan ran in one-shot mode (refine reads ocand / ocount)
  bool emit = false;             // out: the long scan emitted its candidates into ocand (no per-CTA lists)
  int ocap = 0;                  // capacity of ocand per query (kOneshotCap / kEmitCap)
  int n_lists = 0;               // out: per-CTA lists per query
  int kcap = 0;                  // out: capacity of each
};

// One tcgen05 scan pipeline over all of a's queries: prepare the query operand, (seed thresholds on a prefix
// sample,) scan every pass, merge the per-CTA lists.  side != nullptr: filter mode -- the per-CTA lists keep
// kcap(top_k) candidates by upper bound, the merge keeps side->kp of them in side->akeys.  need != nullptr:
// conditional run -- only flagged queries are recomputed and written.  Otherwise exactly one of out_keys /
// (out_scores, out_ids): [Q, top_k].
static int tc_pipeline(Scratch* s, const SearchArgs& a, int mode, FilterSide* side, const uint32_t* need,
                       uint64_t* out_keys, float* out_scores, int64_t* out_ids, cudaStream_t stream) {
  int rc;
  const int kcap = topk_capacity(a.top_k);
  const int list_k = side ? kcap : a.top_k;                 // entries a list keeps
  const bool cached_norms = a.row_sumsq != nullptr;
  int qp = 0, stages = 0;
  bool pair = false, resident = false;
  // Short scans on the filter path: every CTA owns ONE tile, thresholds come from a grid-wide exchange of the tiles'
  // bests, candidates go straight to a global array (scan_tc.cu, oneshot_epilogue).  Needs at least top_k tiles (T is the
  // top_k-th largest of the tiles' bests) and all of them resident at once.
  const int64_t n_tiles_all = (a.n_rows + 127) / 128;
  const bool oneshot = side != nullptr && side->ocand != nullptr && !need && n_tiles_all <= s->sm_count &&
                       n_tiles_all >= a.top_k && a.top_k <= 64 && !(a.flags & (DEVQA_DEBUG_FILTER_ONLY | DEVQA_DEBUG_NO_ONESHOT));
  // Long scans whose candidates refine would otherwise pick out of the per-CTA lists: no lists at all, survivors go
  // straight to the flat candidate array (scan_tc.cu emit_epilogue).  Needs the threshold exchange (>= top_k CTAs per
  // pass; a long scan runs one pass per launch on every SM).
  const bool emit = side != nullptr && side->ocand != nullptr && !oneshot && !need && !side->merge &&
                    s->sm_count >= a.top_k && n_tiles_all >= s->sm_count && !(a.flags & (DEVQA_DEBUG_FILTER_ONLY | DEVQA_DEBUG_NO_EMIT));
  if ((rc = scan_tc_plan(mode, int(a.n_queries), a.dim, kcap, cached_norms, n_tiles_all, s->sm_count, a.flags,
                         &qp, &stages, &pair, oneshot, emit, &resident)))
    return rc;
  const int q_pad = int((a.n_queries + qp - 1) / qp) * qp;
  const size_t qcat_bytes = scan_tc_qcat_bytes(mode, a.dim, q_pad);
  const bool exchange = scan_tc_exchanges_thresholds(mode, a.dim);
  if ((rc = grow(reinterpret_cast<void**>(&s->qsplit), &s->qsplit_bytes,
                 qcat_bytes + size_t(q_pad) * 4 + (exchange && !side ? size_t(q_pad) * kGmaxStride * 4 : 0), false)))
    return rc;
  uint32_t* gthr = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(s->qsplit) + qcat_bytes);   // shared per-query thresholds
  uint32_t* gmax = side ? side->gmax : (exchange ? gthr + q_pad : nullptr);
  const int64_t tiles = (a.n_rows + 127) / 128;
  // Large banks: one pass per launch on every SM.  Small banks: several passes side by side so that
  // tiles x passes still covers the SMs.
  const int n_passes = q_pad / qp;
  const int tiles_cap = int(tiles < s->sm_count ? tiles : s->sm_count);
  int side_by = 1;
  if (tiles < s->sm_count) side_by = n_passes < s->sm_count ? n_passes : s->sm_count;
  int ctas_per_pass = s->sm_count / side_by;
  if (ctas_per_pass > tiles_cap) ctas_per_pass = tiles_cap;
  if (ctas_per_pass < 1) ctas_per_pass = 1;
  if (emit) DEVQA_REQUIRE(ctas_per_pass >= a.top_k, "emit scan: fewer CTAs per pass than top_k");
  if (oneshot) {                                             // one CTA per tile; as many passes per launch as fit
    ctas_per_pass = int(tiles);
    side_by = s->sm_count / ctas_per_pass;
    if (side_by > n_passes) side_by = n_passes;
    DEVQA_REQUIRE(n_passes <= 64, "one-shot scan: too many passes for the arrival counters");
  }
  const int grid = ctas_per_pass;                           // candidate lists per query
  if ((rc = grow(reinterpret_cast<void**>(&s->cand), &s->cand_bytes, size_t(grid) * a.n_queries * kcap * 8, false))) return rc;

  ScanParams p{};
  p.n_rows = a.n_rows; p.dim = a.dim; p.bank_ld = a.bank_ld; p.bank = a.bank;
  p.queries = a.queries; p.query_ld = a.query_ld; p.n_queries = int32_t(a.n_queries);
  p.top_k = list_k; p.kcap = kcap; p.flags = a.flags; p.idmap = a.idmap; p.row_sumsq = a.row_sumsq;
  p.row_scale = a.row_scale; p.qscale = (side && a.row_scale) ? side->qscale : nullptr;
  p.cand = s->cand; p.need = need;
  if (side) {
    p.qerr = side->qerr; p.k_true = a.top_k; p.lostu = side->lostu; p.bank_max_sumsq = a.max_sumsq;
    side->n_lists = grid;
    side->oneshot = oneshot;
    side->emit = emit;
    side->ocap = oneshot ? kOneshotCap : kEmitCap;
    if (oneshot) { p.oneshot = 1; p.gbar = side->gbar; p.ocount = side->ocount; p.ocand = side->ocand; p.ocap = kOneshotCap; }
    if (emit) p.b_resident = resident ? 1 : 0;
    if (emit) { p.emit = 1; p.gbar = side->gbar; p.ocount = side->ocount; p.ocand = side->ocand; p.ocap = kEmitCap; }   // [grid][Q] / [grid][Q][cap]
    if (emit) DEVQA_REQUIRE(n_passes <= 64, "emit scan: too many passes for the arrival counters");
  }
  p.gmax = gmax; p.gmax_stride = kGmaxStride;
  DEVQA_REQUIRE(grid <= kGmaxStride, "more CTAs per pass than the threshold exchange holds");

  const bool prof = s->profiling && !need && s->prof_used < Scratch::kProfSlots;
  if (prof) DEVQA_CUDA(cudaEventRecord(s->prof_ev[2 * s->prof_used], stream));
  if ((rc = scan_tc_prep_queries(mode, a.queries, a.query_ld, int(a.n_queries), qp, q_pad, a.dim, a.flags, s->qsplit, gthr,
                                 side ? side->qinv : nullptr, side ? side->qerr : nullptr, side ? side->seed_u : nullptr,
                                 gmax, kGmaxStride, need, stream, (oneshot || emit) ? side->gbar : nullptr,
                                 oneshot ? 128 + int(a.n_queries) : (emit ? 128 : 0), p.qscale ? side->qscale : nullptr)))
    return rc;
  // Threshold seeding: scan a prefix sample of the bank (one tile per CTA) first; the k-th best score of a
  // SUBSET is a lower bound of the k-th best of the whole bank, so the main scan starts with tight
  // thresholds in every CTA instead of each CTA re-learning them (k ln(n/k) list insertions each).
  // (the filter variant exchanges thresholds between CTAs while it runs and needs no seeding)
  const bool seed = tiles >= 8 * int64_t(ctas_per_pass) && !(a.flags & (1u << 21)) && !need && (!exchange || (a.flags >> 24) & 0xfu);
  if (seed) {
    ScanParams pre = p;
    int seed_tiles = int((a.flags >> 24) & 0xfu);           // undocumented: sample tiles per CTA (benchmark sweeps)
    if (seed_tiles == 0) seed_tiles = 1;
    pre.n_rows = int64_t(ctas_per_pass) * 128 * seed_tiles;
    if (pre.n_rows > a.n_rows) pre.n_rows = a.n_rows;
    for (int pass = 0; pass < n_passes; pass += side_by) {
      const int now = n_passes - pass < side_by ? n_passes - pass : side_by;
      pre.q0 = int32_t(pass * qp);
      pre.nq_pass = qp;
      if ((rc = scan_tc_launch(pre, mode, s->qsplit, gthr, qp, stages, now, ctas_per_pass, false, stream))) return rc;
    }
    // exact modes: the sample's k-th best score IS a threshold; filter: its k-th best upper bound, from which the
    // kernel derives one (minus 2E)
    if ((rc = merge_lists_launch(s->cand, grid, a.n_queries, kcap, kcap, a.top_k, nullptr, nullptr, nullptr, stream,
                                 side ? side->seed_u : gthr)))
      return rc;
    if (side) p.seed_u = side->seed_u;
  }
  for (int pass = 0; pass < n_passes; pass += side_by) {
    const int now = n_passes - pass < side_by ? n_passes - pass : side_by;
    p.q0 = int32_t(pass * qp);
    p.nq_pass = qp;
    if ((rc = scan_tc_launch(p, mode, s->qsplit, gthr, qp, stages, now, ctas_per_pass, pair && now == 1, stream))) return rc;
    if (prof) s->prof_scan_launches++;
  }
  if (prof) {
    DEVQA_CUDA(cudaEventRecord(s->prof_ev[2 * s->prof_used + 1], stream));
    s->prof_used++;
  }
  if (side) {
    side->kcap = kcap;
    if (!side->merge || oneshot || emit) return DEVQA_OK;
    return merge_lists_launch(s->cand, grid, a.n_queries, kcap, kcap, side->kp, side->akeys, nullptr, nullptr, stream);
  }
  return merge_lists_launch(s->cand, grid, a.n_queries, kcap, kcap, a.top_k, out_keys, out_scores, out_ids, stream, nullptr, need);
}

static int run_search(Scratch* s, const SearchArgs& a, uint64_t* out_keys, float* out_scores, int64_t* out_ids,
                      cudaStream_t stream);

// top_k beyond DEVQA_MAX_TOPK (upstream's semantic_search accepts any top_k): results come 1024 at a time, pass j
// being the top-1024 among the keys BELOW the last key pass j - 1 emitted (keys are unique and totally ordered, so the
// passes tile the ranking exactly).  One bank read per query and pass on the stream kernel -- a rarely used corner,
// served for completeness, not for speed.
static int run_search_deep(Scratch* s, const SearchArgs& a, uint64_t* out_keys, float* out_scores, int64_t* out_ids,
                           cudaStream_t stream) {
  int rc;
  const size_t nq = size_t(a.n_queries);
  if ((rc = grow(reinterpret_cast<void**>(&s->deep), &s->deep_bytes, nq * DEVQA_MAX_TOPK * 8 + nq * 8, false))) return rc;
  uint64_t* pass_keys = s->deep;
  uint64_t* below = s->deep + nq * DEVQA_MAX_TOPK;
  for (int col0 = 0; col0 < a.top_k; col0 += DEVQA_MAX_TOPK) {
    SearchArgs aj = a;
    aj.top_k = a.top_k - col0 < DEVQA_MAX_TOPK ? a.top_k - col0 : DEVQA_MAX_TOPK;
    aj.flags |= DEVQA_FORCE_STREAM;
    aj.below = col0 ? below : nullptr;
    if ((rc = run_search(s, aj, pass_keys, nullptr, nullptr, stream))) return rc;
    if ((rc = deep_emit_launch(pass_keys, a.n_queries, aj.top_k, a.top_k, col0, out_keys, out_scores, out_ids, below, stream)))
      return rc;
  }
  return DEVQA_OK;
}

// Exactly one of out_keys / (out_scores,out_ids).
static int run_search(Scratch* s, const SearchArgs& a, uint64_t* out_keys, float* out_scores, int64_t* out_ids,
                      cudaStream_t stream) {
  if (a.n_queries == 0) return DEVQA_OK;
  if (a.top_k > DEVQA_MAX_TOPK) return run_search_deep(s, a, out_keys, out_scores, out_ids, stream);
  const int kcap = topk_capacity(a.top_k);
  const int nv_vec = a.dtype == DEVQA_F32 ? 4 : 8;
  const size_t esz = a.dtype == DEVQA_F32 ? 4 : 2;
  const bool vectorised = (a.dim % nv_vec == 0) && (a.bank_ld % nv_vec == 0) &&
                          (reinterpret_cast<uintptr_t>(a.bank) % 16 == 0);
  const int nv = vectorised ? nv_vec : 1;
  int rc;
  // ---- kernel selection (by batch size only, so that every shard of a sharded bank picks the same one) ----
  // Q <= 2: one pass of the CUDA-core stream (7.0 TB/s on 10M x 768, two launches).  Larger batches: the tcgen05
  // contraction, which streams at the same rate up to 64 queries per pass (the CUDA-core kernel is issue-bound
  // from 4 queries per pass on: 4.87 ms against 4.40 ms sustained on 10M rows).
  const bool tc_ok = scan_tc_supported(a.dtype, a.bank, a.dim, a.bank_ld, a.top_k) && a.n_rows > 0;
  bool use_tc = tc_ok && a.n_queries > 2;
  // Where the tensor path is the filter (see below), it returns bit for bit what the stream kernel returns, so
  // the choice between the two may also look at the bank size without breaking "G shards == one bank":
  const bool filter_path = !(a.flags & DEVQA_FORCE_TF32X3) && a.top_k <= 128 && a.dim <= 16384 &&
                           (a.dtype == DEVQA_F32 ? ((a.flags & DEVQA_NORMALIZE_BANK) || (a.row_sumsq && a.max_sumsq))
                                                 : (a.row_sumsq != nullptr && a.max_sumsq != nullptr));

  // fp16 shadow: where the scan runs one of the epilogues that understand the scaled accumulators -- emit mode (long
  // scans) or one-shot mode (short scans); the conditions mirror tc_pipeline's
  const int64_t sh_tiles = (a.n_rows + 127) / 128;
  const bool sh_emit = sh_tiles >= s->sm_count && s->sm_count >= a.top_k && !(a.flags & DEVQA_DEBUG_NO_EMIT);
  const bool sh_oneshot = sh_tiles <= s->sm_count && sh_tiles >= a.top_k && !(a.flags & DEVQA_DEBUG_NO_ONESHOT);
  const bool shadow_scan = filter_path && a.shadow != nullptr && a.dtype == DEVQA_F32 && a.top_k <= 64 &&
                           (sh_emit || sh_oneshot) && !(a.flags & (DEVQA_DEBUG_FILTER_ONLY | DEVQA_NO_SHADOW));
  if (tc_ok && filter_path) {
    const int64_t tiles = (a.n_rows + 127) / 128;
    // tiny banks (fewer tiles than SMs), up to 8 queries: ONE stream launch (~15 us) beats the seven-launch pipeline
    if (a.n_queries <= 8 && tiles < s->sm_count) use_tc = false;
    // large bf16 banks: the bf16 stream kernel manages 5.1 TB/s, the tensor path 6.9 -- also for one or two queries
    if ((a.dtype == DEVQA_BF16 || shadow_scan) && a.n_queries <= 2 && tiles >= 8 * int64_t(s->sm_count)) use_tc = true;
  }
  if (a.flags & DEVQA_FORCE_STREAM) use_tc = false;
  if (a.flags & DEVQA_FORCE_TENSOR) {
    if (!tc_ok) {
      set_error("DEVQA_FORCE_TENSOR: the tcgen05 scan needs 16-byte aligned bank rows and top_k <= 128");
      return DEVQA_E_UNSUPPORTED;
    }
    use_tc = true;
  }

  ScanParams p{};
  p.n_rows = a.n_rows; p.dim = a.dim; p.bank_ld = a.bank_ld;
  p.queries = a.queries; p.query_ld = a.query_ld; p.n_queries = int32_t(a.n_queries);
  p.top_k = a.top_k; p.kcap = kcap; p.flags = a.flags; p.idmap = a.idmap; p.row_sumsq = a.row_sumsq;
  p.below = a.below;
  int grid = 1;

  if (use_tc) {
    // fp32 banks: the TF32 filter scan + exact re-scoring (refine.cu) + conditional exact fallback; the 3xTF32
    // scan alone on request, for top_k > 64 (no room for the filter's surplus candidates) and for un-normalised
    // scores over a caller-owned bank (the error bound needs max ||row||, which only library-owned banks cache).
    // K': candidates kept per query after the merge
    const int kp = a.top_k <= 8 ? 32 : (a.top_k <= 16 ? 64 : (a.top_k <= 64 ? 128 : 256));
    // (Nothing here may depend on the number of rows: every shard of a sharded bank has to take the same path, or
    // G-GPU results would differ from 1-GPU results in the last bits of the scores.)
    // bf16 banks: the same filter with queries rounded to bf16 (library-owned banks: cached row norms), bf16x3 as fallback.
    const bool filt = filter_path;
    const int filt_mode = (a.dtype == DEVQA_F32 && !shadow_scan) ? kTcTf32x1 : kTcBf16x1;
    const int exact_mode = a.dtype == DEVQA_F32 ? kTcTf32x3 : kTcBf16x3;
    if (s->profiling) s->prof_searches++;
    constexpr int64_t kChunkQ = 1024;                                  // bounds the candidate scratch (grid x Q x kcap keys)
    for (int64_t c0 = 0, clen = 0; c0 < a.n_queries; c0 += clen) {
      SearchArgs ac = a;
      ac.queries = a.queries + c0 * a.query_ld;
      ac.n_queries = a.n_queries - c0 < kChunkQ ? a.n_queries - c0 : kChunkQ;
      // 256-query passes: a ragged tail (e.g. 300 = 256 + 44) goes into its own chunk, which then plans a narrower,
      // faster pass for itself instead of a second 256-wide one that is mostly padding
      // (a tail of up to 128 queries costs ~5 ms on 10M rows as its own pass, ~7 ms as a padded 256-wide one)
      if (filt && a.row_sumsq && ac.n_queries > 256 && ac.n_queries % 256 != 0 && ac.n_queries % 256 <= 128)
        ac.n_queries -= ac.n_queries % 256;
      clen = ac.n_queries;
      uint64_t* ok = out_keys ? out_keys + c0 * a.top_k : nullptr;
      float* os = out_scores ? out_scores + c0 * a.top_k : nullptr;
      int64_t* oi = out_ids ? out_ids + c0 * a.top_k : nullptr;
      if (!filt) {
        if ((rc = tc_pipeline(s, ac, a.dtype == DEVQA_F32 ? kTcTf32x3 : kTcBf16x3, nullptr, nullptr, ok, os, oi, stream)))
          return rc;
        continue;
      }
      const size_t nq = size_t(ac.n_queries), nq_pad = (nq + 255) / 256 * 256;   // >= q_pad of any plan
      if ((rc = grow(reinterpret_cast<void**>(&s->akeys), &s->akeys_bytes,
                     nq * kp * 8 + nq * 16 + nq_pad * 4 + size_t(s->sm_count) * nq * 4 + nq_pad * kGmaxStride * 4 +
                         (128 + size_t(s->sm_count) * nq_pad) * 4 + 8 +
                         nq * (size_t(kOneshotCap) > size_t(s->sm_count) * kEmitCap ? size_t(kOneshotCap) : size_t(s->sm_count) * kEmitCap) * 8,
                     false)))
        return rc;
      FilterSide side;
      side.kp = kp;
      side.akeys = s->akeys;
      side.need = reinterpret_cast<uint32_t*>(s->akeys + nq * kp);
      side.qinv = reinterpret_cast<float*>(side.need + nq);
      side.qerr = side.qinv + nq;
      side.qscale = side.qerr + nq;
      side.seed_u = reinterpret_cast<uint32_t*>(side.qscale + nq);
      side.lostu = side.seed_u + nq_pad;
      side.gmax = side.lostu + size_t(s->sm_count) * nq;
      side.gbar = side.gmax + nq_pad * kGmaxStride;                     // [128] counters, then [nq] candidate counts,
      side.ocount = side.gbar + 128;                                    // zeroed as ONE region by the prep kernel
      side.ocand = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(side.ocount + size_t(s->sm_count) * nq_pad) + 7) & ~uintptr_t(7));
      // With at least top_k CTAs per pass the scan's CTAs exchange their bests, and refine can pick the candidates
      // straight out of the per-CTA lists (one kernel instead of merge + refine).
      const int64_t tiles = (a.n_rows + 127) / 128;
      // (top_k > 64: the top_k-th largest of <= 148 CTA bests is too loose a cut for refine_lists -- merged path, K' = 256)
      side.merge = (a.flags & DEVQA_DEBUG_FILTER_ONLY) || tiles < s->sm_count || s->sm_count < a.top_k || a.top_k > 64;
      {
        SearchArgs as = ac;                                              // what the filter SCANS
        if (shadow_scan) {
          as.bank = a.shadow; as.bank_ld = a.shadow_ld; as.dtype = DEVQA_BF16 /* 2-byte rows */; as.flags |= kFlagShadowScan;
          as.row_scale = a.shadow_scale;
        }
        if ((rc = tc_pipeline(s, as, filt_mode, &side, nullptr, nullptr, nullptr, nullptr, stream))) return rc;
      }
      if (a.flags & DEVQA_DEBUG_FILTER_ONLY) {                           // tests: the filter's own upper bounds
        if ((rc = merge_lists_launch(side.akeys, 1, ac.n_queries, kp, kp, a.top_k, ok, os, oi, stream))) return rc;
        continue;
      }
      RefineParams rp{};
      rp.bank = a.bank; rp.bank_dtype = a.dtype; rp.n_rows = a.n_rows; rp.dim = a.dim; rp.bank_ld = a.bank_ld;
      rp.queries = ac.queries; rp.query_ld = a.query_ld; rp.n_queries = int32_t(ac.n_queries);
      rp.qinv = side.qinv; rp.akeys = side.akeys; rp.lostu = side.lostu; rp.n_lists = side.n_lists;
      rp.kp = kp; rp.top_k = a.top_k; rp.flags = a.flags; rp.idmap = a.idmap;
      rp.need = side.need; rp.fallbacks = s->fallbacks; rp.violations = s->violations; rp.out_keys = ok; rp.out_scores = os; rp.out_ids = oi;
      rp.cand = s->cand; rp.cand_kcap = side.kcap; rp.gmax = side.gmax; rp.gmax_stride = kGmaxStride;
      rp.qerr = side.qerr; rp.bank_max_sumsq = a.max_sumsq;
      if (side.oneshot) { rp.dcount = side.ocount; rp.dcand = side.ocand; rp.dcap = side.ocap; }
      if (side.emit) { rp.cand = side.ocand; rp.cand_kcap = kEmitCap; rp.cand_count = side.ocount; }   // per-CTA regions
      if ((rc = (side.merge && !side.oneshot) ? refine_launch(rp, stream) : refine_lists_launch(rp, stream))) return rc;
      // queries for which a lost row could still matter are redone exactly; for all others these launches
      // exit at once (the flags live on the device: no host round trip, graph-capturable).  The fallback's top-k
      // (3xTF32 scores) is then re-scored in the stream kernel's arithmetic like every other result, so scores
      // stay bit-identical across kernels, batch sizes and shardings even for flagged queries.
      // The fallback hands over a few rows MORE than top_k: two rows whose scores differ by less than the rounding
      // difference between 3xTF32 and the stream kernel's fp32 sums may swap places across the cut, and the final
      // selection has to be the stream kernel's.
      uint64_t* fkeys = side.akeys;                                      // [Q, fk] keys; akeys is free again
      SearchArgs af = ac;
      int fk = a.top_k + 8;
      if (fk > 128) fk = 128;
      if (fk > kp) fk = kp;
      if (int64_t(fk) > a.n_rows) fk = int(a.n_rows);
      if (fk < a.top_k) fk = a.top_k;
      af.top_k = fk;
      if ((rc = tc_pipeline(s, af, exact_mode, nullptr, side.need, fkeys, nullptr, nullptr, stream))) return rc;
      RefineParams fr = rp;
      fr.akeys = fkeys; fr.kp = fk; fr.only_if = side.need; fr.lostu = nullptr; fr.n_lists = 0;
      if ((rc = refine_launch(fr, stream))) return rc;
    }
    return DEVQA_OK;
  }

  // widest query tile whose shared-memory footprint fits
  int qt = 8;
  while (qt > 1 && (qt / 2 >= a.n_queries || scan_stream_smem_bytes(a.dim, nv, qt, kcap) > 200 * 1024)) qt >>= 1;
  if (scan_stream_smem_bytes(a.dim, nv, qt, kcap) > 220 * 1024) {
    set_error("dim %lld with top_k %d does not fit the scan kernel's shared memory", (long long)a.dim, a.top_k);
    return DEVQA_E_UNSUPPORTED;
  }
  const int R = scan_stream_rows_per_group(qt);
  const int64_t groups = (a.n_rows + R - 1) / R;
  int64_t want_ctas = (groups + (kScanThreads / 32) - 1) / (kScanThreads / 32);
  grid = int(want_ctas < s->sm_count ? want_ctas : s->sm_count);
  if (grid < 1) grid = 1;

  const size_t cand_bytes = size_t(grid) * a.n_queries * kcap * 8;
  if ((rc = grow(reinterpret_cast<void**>(&s->cand), &s->cand_bytes, cand_bytes, false))) return rc;

  const void* bank_ptr = a.bank;
  if (a.n_rows > 0 && a.n_rows < 8) {
    // the scan kernel reads whole groups of up to 8 rows: give it a zero-padded copy (rows >= n_rows are masked)
    const size_t row_bytes = size_t(a.bank_ld) * esz;
    if ((rc = grow(&s->tiny, &s->tiny_bytes, 8 * row_bytes, false))) return rc;
    DEVQA_CUDA(cudaMemsetAsync(s->tiny, 0, 8 * row_bytes, stream));
    DEVQA_CUDA(cudaMemcpyAsync(s->tiny, a.bank, size_t(a.n_rows - 1) * row_bytes + size_t(a.dim) * esz,
                               cudaMemcpyDeviceToDevice, stream));
    bank_ptr = s->tiny;
  }
  p.bank = bank_ptr; p.cand = s->cand;
  // the last CTA of each pass merges: no merge launch (DEVQA_FUSED_MERGE=0 restores the separate merge kernel)
  static const bool fused_ok = getenv("DEVQA_FUSED_MERGE") == nullptr || atoi(getenv("DEVQA_FUSED_MERGE")) != 0;
  const bool fused_merge = fused_ok && kcap == 32;
  if (fused_merge) { p.ticket = s->ticket; p.out_keys = out_keys; p.out_scores = out_scores; p.out_ids = out_ids; }
  const bool prof = s->profiling && s->prof_used < Scratch::kProfSlots;
  if (s->profiling) s->prof_searches++;
  if (prof) DEVQA_CUDA(cudaEventRecord(s->prof_ev[2 * s->prof_used], stream));
  for (int64_t q0 = 0; q0 < a.n_queries; q0 += qt) {
    p.q0 = int32_t(q0);
    p.nq_pass = int32_t(a.n_queries - q0 < qt ? a.n_queries - q0 : qt);
    rc = scan_stream_launch(p, a.dtype, vectorised, qt, grid, stream);
    if (rc) return rc;
    if (prof) s->prof_scan_launches++;
  }
  if (prof) {
    DEVQA_CUDA(cudaEventRecord(s->prof_ev[2 * s->prof_used + 1], stream));
    s->prof_used++;
  }
  if (fused_merge) return DEVQA_OK;
  return merge_lists_launch(s->cand, grid, a.n_queries, kcap, kcap, a.top_k, out_keys, out_scores, out_ids, stream);
}

// ---- bank ---------------------------------------------------------------------------------------------
}  // namespace devqa

struct devqa_bank {
  int device = 0;
  int dtype = DEVQA_F32;
  int64_t dim = 0, ld = 0;
  int64_t size = 0, capacity = 0;
  void* data = nullptr;
  devqa::VmmBuffer* vmm = nullptr; // rows live in a reserved virtual range that grows in place (vmm.cu); nullptr: cudaMalloc + copy
  // Optional 2-byte (scaled fp16) SHADOW of an fp32 bank (devqa_bank_set_shadow): the tensor-core filter streams these 2 bytes per
  // element instead of 4, candidates are re-scored against the fp32 rows -- exact fp32 results from half the bytes.
  void* shadow = nullptr;
  devqa::VmmBuffer* shadow_vmm = nullptr;
  float* shadow_scale = nullptr;   // per row: 2^-e of the power-of-two scaling its shadow row was stored with
  int64_t shadow_ld = 0, shadow_capacity = 0;
  float* sumsq = nullptr;          // cached ||row||^2 per stored row (bf16 banks: of the rounded row)
  uint32_t* max_sumsq = nullptr;   // device word: max of those over every row ever appended (float bits; never shrinks)
  devqa::IdMap idmap{0, 0, 0, 1};
  size_t esz() const { return dtype == DEVQA_F32 ? 4 : 2; }
};

struct devqa_comm {
  int rank = 0, world = 1, device = 0;
  void* nccl = nullptr;   // ncclComm_t
  // Peer-to-peer exchange of the local top-k (scan.cuh, P2pExchange).  Mailbox of a rank, in its own memory:
  //   [0, 256)                      header (word 0: waits that timed out)
  //   flags  [2 slots][world][kP2pMaxQ]  u32    sequence number of the call whose keys have arrived
  //   data   [2 slots][world][kP2pMaxKeys] u64  the keys, per slot laid out [source rank][n_queries][top_k]
  unsigned char* mbox = nullptr;                             // this rank's mailbox (cudaMalloc, exported by IPC handle)
  unsigned char* peer_mbox[devqa::kP2pMaxWorld] = {};        // every rank's mailbox as mapped into this process
  bool p2p = false;
  size_t flags_off() const { return 256; }
  size_t data_off() const { return 256 + size_t(2) * world * devqa::kP2pMaxQ * 4; }
  size_t mbox_bytes() const { return data_off() + size_t(2) * world * devqa::kP2pMaxKeys * 8; }
};

namespace devqa {

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define DEVQA_ON_DEVICE(dev)                                       \
  ::devqa::DeviceGuard guard__(dev);                               \
  if (!guard__.ok) { ::devqa::set_error("cudaSetDevice(%d) failed", dev); return DEVQA_E_CUDA; }

static int bank_grow(devqa_bank* b, int64_t want_rows, cudaStream_t stream) {
  if (want_rows <= b->capacity) return DEVQA_OK;
  int64_t cap = b->capacity > 0 ? b->capacity : 1024;
  while (cap < want_rows) cap += cap / 2 + 1024;          // geometric: amortised O(1) per appended row
  const size_t bytes = size_t(cap) * b->ld * b->esz();
  // the row norms (4 bytes per row) are small enough to move
  float* fresh_ss = nullptr;
  cudaError_t e = cudaMalloc(&fresh_ss, size_t(cap) * 4);
  if (e != cudaSuccess) {
    set_error("cudaMalloc of %zu bytes for the row norms failed: %s", size_t(cap) * 4, cudaGetErrorString(e));
    return DEVQA_E_NOMEM;
  }
  if (b->vmm) {
    // in place: fresh physical chunks are mapped behind the rows already stored -- nothing is copied, nothing moves
    bool moved = false;
    int rc = vmm_grow(b->vmm, bytes, &moved);
    if (rc == DEVQA_OK) {
      if (b->size > 0) DEVQA_CUDA(cudaMemcpyAsync(fresh_ss, b->sumsq, size_t(b->size) * 4, cudaMemcpyDeviceToDevice, stream));
      if (b->sumsq) { DEVQA_CUDA(cudaStreamSynchronize(stream)); DEVQA_CUDA(cudaFree(b->sumsq)); }
      b->data = vmm_data(b->vmm);
      b->sumsq = fresh_ss;
      b->capacity = cap;
      return DEVQA_OK;
    }
    if (rc != DEVQA_E_UNSUPPORTED || b->data != nullptr) { cudaFree(fresh_ss); return rc; }
    vmm_destroy(b->vmm);                                  // (never worked on this device: plain allocations from here on)
    b->vmm = nullptr;
  }
  void* fresh = nullptr;
  e = cudaMalloc(&fresh, bytes);
  if (e != cudaSuccess) {
    cudaFree(fresh_ss);
    set_error("cudaMalloc of %zu bytes for the bank failed: %s", bytes, cudaGetErrorString(e));
    return DEVQA_E_NOMEM;
  }
  if (b->size > 0) {
    DEVQA_CUDA(cudaMemcpyAsync(fresh, b->data, size_t(b->size) * b->ld * b->esz(), cudaMemcpyDeviceToDevice, stream));
    DEVQA_CUDA(cudaMemcpyAsync(fresh_ss, b->sumsq, size_t(b->size) * 4, cudaMemcpyDeviceToDevice, stream));
  }
  if (b->data) {
    // the old buffers may still be read by searches enqueued on ANY stream: wait for the device, not just this stream
    DEVQA_CUDA(cudaDeviceSynchronize());
    DEVQA_CUDA(cudaFree(b->data));
    DEVQA_CUDA(cudaFree(b->sumsq));
  }
  b->data = fresh;
  b->sumsq = fresh_ss;
  b->capacity = cap;
  return DEVQA_OK;
}

// keep the shadow's storage in step with the bank's capacity
static int shadow_grow(devqa_bank* b, int64_t cap, cudaStream_t stream) {
  if (!b->shadow_ld || cap <= b->shadow_capacity) return DEVQA_OK;
  const size_t bytes = size_t(cap) * b->shadow_ld * 2;
  {                                                          // the per-row scales (4 bytes per row) are small enough to move
    float* fresh_sc = nullptr;
    cudaError_t e = cudaMalloc(&fresh_sc, size_t(cap) * 4);
    if (e != cudaSuccess) {
      set_error("cudaMalloc of %zu bytes for the shadow's row scales failed: %s", size_t(cap) * 4, cudaGetErrorString(e));
      return DEVQA_E_NOMEM;
    }
    if (b->shadow_scale) {
      if (b->size > 0) DEVQA_CUDA(cudaMemcpyAsync(fresh_sc, b->shadow_scale, size_t(b->size) * 4, cudaMemcpyDeviceToDevice, stream));
      DEVQA_CUDA(cudaDeviceSynchronize());
      DEVQA_CUDA(cudaFree(b->shadow_scale));
    }
    b->shadow_scale = fresh_sc;
  }
  if (b->shadow_vmm) {
    bool moved = false;
    int rc = vmm_grow(b->shadow_vmm, bytes, &moved);
    if (rc == DEVQA_OK) { b->shadow = vmm_data(b->shadow_vmm); b->shadow_capacity = cap; return DEVQA_OK; }
    if (rc != DEVQA_E_UNSUPPORTED || b->shadow != nullptr) return rc;
    vmm_destroy(b->shadow_vmm);
    b->shadow_vmm = nullptr;
  }
  void* fresh = nullptr;
  cudaError_t e = cudaMalloc(&fresh, bytes);
  if (e != cudaSuccess) {
    set_error("cudaMalloc of %zu bytes for the shadow failed: %s", bytes, cudaGetErrorString(e));
    return DEVQA_E_NOMEM;
  }
  if (b->shadow) {
    if (b->size > 0)
      DEVQA_CUDA(cudaMemcpyAsync(fresh, b->shadow, size_t(b->size) * b->shadow_ld * 2, cudaMemcpyDeviceToDevice, stream));
    DEVQA_CUDA(cudaDeviceSynchronize());
    DEVQA_CUDA(cudaFree(b->shadow));
  }
  b->shadow = fresh;
  b->shadow_capacity = cap;
  return DEVQA_OK;
}

// rows [first, first + n) were just written to the fp32 bank: round them into the shadow
static int shadow_update(devqa_bank* b, int64_t first, int64_t n, cudaStream_t stream) {
  if (!b->shadow_ld || n == 0) return DEVQA_OK;
  int rc = shadow_grow(b, b->capacity, stream);
  if (rc) return rc;
  const float* src = static_cast<const float*>(b->data) + size_t(first) * b->ld;
  char* dst = static_cast<char*>(b->shadow) + size_t(first) * b->shadow_ld * 2;
  return shadow_convert_launch(src, b->ld, dst, b->shadow_ld, b->shadow_scale + first, n, b->dim, stream);
}

static SearchArgs bank_args(const devqa_bank* b, const float* q, int64_t nq, int64_t qld, int k, unsigned flags) {
  SearchArgs a{};
  a.bank = b->data; a.dtype = b->dtype; a.n_rows = b->size; a.dim = b->dim; a.bank_ld = b->ld;
  a.queries = q; a.n_queries = nq; a.query_ld = qld; a.top_k = k; a.flags = flags; a.idmap = b->idmap;
  a.row_sumsq = b->sumsq;
  a.max_sumsq = b->max_sumsq;
  a.shadow = b->shadow; a.shadow_ld = b->shadow_ld; a.shadow_scale = b->shadow_scale;
  return a;
}

// H2D of queries into scratch; returns device pointer
// (tried: results up to 64 KB straight into pinned memory for any query batch -- end to end 26.2 k -> 23.2 k q/s at
// Q = 64: 24 KB of scattered 4/8-byte stores over PCIe cost more than one D2H copy)
static bool small_io(size_t qbytes, size_t obytes) { return qbytes <= 16 * 1024 && obytes <= 16 * 1024; }

static int stage_queries(Scratch* s, const float* qh, int64_t nq, int64_t dim, int64_t qld, int top_k,
                         float** dq, float** d_scores, int64_t** d_ids, cudaStream_t stream) {
  const size_t qbytes = size_t(nq) * dim * 4;
  const size_t obytes = size_t(nq) * top_k * 12;
  const size_t total = ((qbytes + 255) / 256) * 256 + obytes + 256;
  int rc = grow(&s->dev_io, &s->dev_io_bytes, total, false);
  if (rc) return rc;
  rc = grow(&s->pin_io, &s->pin_io_bytes, total, true);
  if (rc) return rc;
  float* pq = static_cast<float*>(s->pin_io);
  for (int64_t i = 0; i < nq; ++i) memcpy(pq + i * dim, qh + i * qld, size_t(dim) * 4);
  // Small calls (the reference's one query at a time): the merge kernel writes the few result bytes straight into
  // the pinned staging buffer (device-accessible under unified addressing) -- one copy command less per search.
  // (Reading the QUERIES from host memory was tried too: every CTA fetching them across the bus doubled the
  // end-to-end latency, 50 -> 104 us at 6345 x 384.)
  *dq = static_cast<float*>(s->dev_io);
  char* base = small_io(qbytes, obytes) ? static_cast<char*>(s->pin_io) : static_cast<char*>(s->dev_io);
  char* outs = base + ((qbytes + 255) / 256) * 256;
  *d_ids = reinterpret_cast<int64_t*>(outs);
  *d_scores = reinterpret_cast<float*>(outs + size_t(nq) * top_k * 8);
  DEVQA_CUDA(cudaMemcpyAsync(*dq, pq, qbytes, cudaMemcpyHostToDevice, stream));
  return DEVQA_OK;
}
static int unstage_results(Scratch* s, int64_t nq, int64_t dim, int top_k, float* scores_h, int64_t* ids_h,
                           cudaStream_t stream) {
  const size_t qbytes = size_t(nq) * dim * 4;
  const size_t off = ((qbytes + 255) / 256) * 256;
  const size_t obytes = size_t(nq) * top_k * 12;
  char* pin_out = static_cast<char*>(s->pin_io) + off;
  if (!small_io(qbytes, obytes))
    DEVQA_CUDA(cudaMemcpyAsync(pin_out, static_cast<char*>(s->dev_io) + off, obytes, cudaMemcpyDeviceToHost, stream));
  DEVQA_CUDA(cudaStreamSynchronize(stream));
  memcpy(ids_h, pin_out, size_t(nq) * top_k * 8);
  memcpy(scores_h, pin_out + size_t(nq) * top_k * 8, size_t(nq) * top_k * 4);
  return DEVQA_OK;
}

}  // namespace devqa

using namespace devqa;

extern "C" {

int devqa_version(void) { return DEVQA_VERSION; }
const char* devqa_last_error(void) { return g_err; }
int64_t devqa_launch_count(void) { return g_launches.load(); }

int devqa_debug_tc_trace(int64_t* device_buf_16x256) {
  scan_tc_set_trace(reinterpret_cast<long long*>(device_buf_16x256));
  return DEVQA_OK;
}

int devqa_filter_fallbacks(int device, int64_t* count) {
  DEVQA_REQUIRE(count != nullptr, "count is null");
  Scratch* s = nullptr;
  int rc = scratch_for(device, &s);
  if (rc) return rc;
  DEVQA_ON_DEVICE(device);
  std::lock_guard<std::mutex> lock(s->mu);
  uint32_t v = 0;
  DEVQA_CUDA(cudaDeviceSynchronize());
  DEVQA_CUDA(cudaMemcpy(&v, s->fallbacks, 4, cudaMemcpyDeviceToHost));
  DEVQA_CUDA(cudaMemset(s->fallbacks, 0, 4));
  *count = v;
  return DEVQA_OK;
}

int devqa_filter_bound_violations(int device, int64_t* count) {
  DEVQA_REQUIRE(count != nullptr, "count is null");
  Scratch* s = nullptr;
  int rc = scratch_for(device, &s);
  if (rc) return rc;
  DEVQA_ON_DEVICE(device);
  std::lock_guard<std::mutex> lock(s->mu);
  uint32_t v = 0;
  DEVQA_CUDA(cudaDeviceSynchronize());
  DEVQA_CUDA(cudaMemcpy(&v, s->violations, 4, cudaMemcpyDeviceToHost));
  DEVQA_CUDA(cudaMemset(s->violations, 0, 4));
  *count = v;
  return DEVQA_OK;
}

double devqa_filter_relerr(int64_t dim, int bank_dtype) { return double(tc_filter_relerr(dim, bank_dtype == DEVQA_BF16)); }

int devqa_profile_enable(int device, int enable) {
  Scratch* s = nullptr;
  int rc = scratch_for(device, &s);
  if (rc) return rc;
  DEVQA_ON_DEVICE(device);
  std::lock_guard<std::mutex> lock(s->mu);
  if (enable && !s->prof_ev) {
    s->prof_ev = new (std::nothrow) cudaEvent_t[2 * Scratch::kProfSlots];
    if (!s->prof_ev) { set_error("host allocation failed"); return DEVQA_E_NOMEM; }
    for (int i = 0; i < 2 * Scratch::kProfSlots; ++i) DEVQA_CUDA(cudaEventCreate(&s->prof_ev[i]));
  }
  s->profiling = enable != 0;
  s->prof_used = 0;
  s->prof_scan_launches = 0;
  s->prof_searches = 0;
  return DEVQA_OK;
}

int devqa_profile_collect(int device, double* scan_ms_total, int64_t* scan_launches, int64_t* searches) {
  Scratch* s = nullptr;
  int rc = scratch_for(device, &s);
  if (rc) return rc;
  DEVQA_REQUIRE(scan_ms_total && scan_launches && searches, "null outputs");
  DEVQA_ON_DEVICE(device);
  std::lock_guard<std::mutex> lock(s->mu);
  double total = 0.0;
  for (int i = 0; i < s->prof_used; ++i) {
    DEVQA_CUDA(cudaEventSynchronize(s->prof_ev[2 * i + 1]));
    float ms = 0.f;
    DEVQA_CUDA(cudaEventElapsedTime(&ms, s->prof_ev[2 * i], s->prof_ev[2 * i + 1]));
    total += ms;
  }
  *scan_ms_total = total;
  *scan_launches = s->prof_scan_launches;
  *searches = s->prof_searches;
  s->prof_used = 0;
  s->prof_scan_launches = 0;
  s->prof_searches = 0;
  return DEVQA_OK;
}

static int search_raw_common(const void* bank, int bank_dtype, int64_t n_rows, int64_t dim, int64_t bank_ld,
                             const float* queries, int64_t n_queries, int64_t query_ld, int top_k, unsigned flags,
                             int64_t id_base, const float* row_sumsq, const uint32_t* max_sumsq_bits,
                             float* out_scores, int64_t* out_ids, void* stream_, const void* shadow = nullptr,
                             int64_t shadow_ld = 0, const float* shadow_scale = nullptr) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SearchArgs a{bank, bank_dtype, n_rows, dim, bank_ld, queries, n_queries, query_ld, top_k, flags,
               IdMap{id_base, 0, 0, 1}};
  // fp32 banks only: for bf16 banks cached norms also enter the exact variants' final scores, and the stream
  // kernel would not use them -- results must not depend on which kernel served the call
  if (row_sumsq && max_sumsq_bits && bank_dtype == DEVQA_F32) { a.row_sumsq = row_sumsq; a.max_sumsq = max_sumsq_bits; }
  if (shadow && shadow_scale && a.row_sumsq) {               // (the filter over the shadow needs the norms as well)
    DEVQA_REQUIRE(bank_dtype == DEVQA_F32, "a shadow only makes sense for fp32 rows");
    DEVQA_REQUIRE(shadow_ld >= dim && shadow_ld % 8 == 0 && reinterpret_cast<uintptr_t>(shadow) % 16 == 0,
                  "shadow rows must be 16-byte aligned (shadow_ld a multiple of 8)");
    a.shadow = shadow; a.shadow_ld = shadow_ld; a.shadow_scale = shadow_scale;
  }
  int rc = validate(a);
  if (rc) return rc;
  DEVQA_REQUIRE(n_queries == 0 || (out_scores && out_ids), "output pointers are null");
  int device = 0;
  DEVQA_CUDA(cudaGetDevice(&device));
  Scratch* s = nullptr;
  rc = scratch_for(device, &s);
  if (rc) return rc;
  std::lock_guard<std::mutex> lock(s->mu);
  if ((rc = scratch_enter(s, stream))) return rc;
  if ((rc = run_search(s, a, nullptr, out_scores, out_ids, stream))) return rc;
  return scratch_leave(s, stream);
}

int devqa_search_raw(const void* bank, int bank_dtype, int64_t n_rows, int64_t dim, int64_t bank_ld,
                     const float* queries, int64_t n_queries, int64_t query_ld, int top_k, unsigned flags,
                     int64_t id_base, float* out_scores, int64_t* out_ids, void* stream_) {
  return search_raw_common(bank, bank_dtype, n_rows, dim, bank_ld, queries, n_queries, query_ld, top_k, flags, id_base,
                           nullptr, nullptr, out_scores, out_ids, stream_);
}

int devqa_search_raw_ex(const void* bank, int bank_dtype, int64_t n_rows, int64_t dim, int64_t bank_ld,
                        const float* queries, int64_t n_queries, int64_t query_ld, int top_k, unsigned flags,
                        int64_t id_base, const float* row_sumsq, const uint32_t* max_sumsq_bits, float* out_scores,
                        int64_t* out_ids, void* stream_) {
  return search_raw_common(bank, bank_dtype, n_rows, dim, bank_ld, queries, n_queries, query_ld, top_k, flags, id_base,
                           row_sumsq, max_sumsq_bits, out_scores, out_ids, stream_);
}

int devqa_search_raw_shadowed(const void* bank, int bank_dtype, int64_t n_rows, int64_t dim, int64_t bank_ld,
                              const float* queries, int64_t n_queries, int64_t query_ld, int top_k, unsigned flags,
                              int64_t id_base, const float* row_sumsq, const uint32_t* max_sumsq_bits, const void* shadow,
                              int64_t shadow_ld, const float* shadow_scale, float* out_scores, int64_t* out_ids,
                              void* stream_) {
  return search_raw_common(bank, bank_dtype, n_rows, dim, bank_ld, queries, n_queries, query_ld, top_k, flags, id_base,
                           row_sumsq, max_sumsq_bits, out_scores, out_ids, stream_, shadow, shadow_ld, shadow_scale);
}

int devqa_make_shadow(const float* bank, int64_t n_rows, int64_t dim, int64_t bank_ld, void* shadow, int64_t shadow_ld,
                      float* shadow_scale, void* stream_) {
  DEVQA_REQUIRE(n_rows >= 0 && dim > 0 && bank_ld >= dim, "bad shape");
  DEVQA_REQUIRE(shadow_ld >= dim && shadow_ld % 8 == 0, "shadow_ld must be >= dim and a multiple of 8 (16-byte rows)");
  DEVQA_REQUIRE(n_rows == 0 || (bank && shadow && shadow_scale), "null pointers");
  if (n_rows == 0) return DEVQA_OK;
  return shadow_convert_launch(bank, bank_ld, shadow, shadow_ld, shadow_scale, n_rows, dim, static_cast<cudaStream_t>(stream_));
}

int devqa_row_sumsq(const void* bank, int bank_dtype, int64_t n_rows, int64_t dim, int64_t bank_ld, float* out_sumsq,
                    uint32_t* out_max_bits, void* stream_) {
  DEVQA_REQUIRE(n_rows >= 0 && dim > 0 && bank_ld >= dim, "bad shape");
  DEVQA_REQUIRE(bank_dtype == DEVQA_F32 || bank_dtype == DEVQA_BF16, "unknown bank dtype %d", bank_dtype);
  DEVQA_REQUIRE(n_rows == 0 || (bank && out_sumsq), "null pointers");
  return row_sumsq_launch(bank, bank_dtype, n_rows, dim, bank_ld, out_sumsq, static_cast<cudaStream_t>(stream_), out_max_bits);
}

int devqa_normalize_rows(const float* src, int64_t src_ld, float* dst, int64_t dst_ld, int64_t n_rows, int64_t dim,
                         unsigned flags, void* stream) {
  DEVQA_REQUIRE(n_rows >= 0 && dim > 0 && src_ld >= dim && dst_ld >= dim, "bad shape");
  DEVQA_REQUIRE(n_rows == 0 || (src && dst), "null pointers");
  return normalize_rows_launch(src, src_ld, dst, dst_ld, n_rows, dim, flags, static_cast<cudaStream_t>(stream));
}

int devqa_bank_create(devqa_bank** out, int device, int64_t dim, int64_t capacity_rows, int dtype) {
  DEVQA_REQUIRE(out != nullptr, "out is null");
  DEVQA_REQUIRE(dim > 0 && dim <= (1 << 20), "dim %lld unsupported", (long long)dim);
  DEVQA_REQUIRE(dtype == DEVQA_F32 || dtype == DEVQA_BF16, "unknown dtype %d", dtype);
  DEVQA_REQUIRE(capacity_rows >= 0, "negative capacity");
  Scratch* s = nullptr;
  int rc = scratch_for(device, &s);
  if (rc) return rc;
  DEVQA_ON_DEVICE(device);
  devqa_bank* b = new (std::nothrow) devqa_bank();
  if (!b) { set_error("host allocation failed"); return DEVQA_E_NOMEM; }
  b->device = device; b->dtype = dtype; b->dim = dim;
  const int64_t align = dtype == DEVQA_F32 ? 4 : 8;          // rows stay 16-byte aligned
  b->ld = (dim + align - 1) / align * align;
  if (cudaMalloc(&b->max_sumsq, 4) != cudaSuccess || cudaMemset(b->max_sumsq, 0, 4) != cudaSuccess) {
    set_error("cudaMalloc for the bank header failed");
    delete b;
    return DEVQA_E_NOMEM;
  }
  static const bool vmm_on = getenv("DEVQA_VMM") == nullptr || atoi(getenv("DEVQA_VMM")) != 0;
  if (vmm_on) b->vmm = vmm_create(device);                   // nullptr: driver without VMM -> cudaMalloc + copy-on-regrow
  if (capacity_rows > 0 && (rc = bank_grow(b, capacity_rows, nullptr))) {
    cudaFree(b->max_sumsq);
    if (b->vmm) vmm_destroy(b->vmm);
    delete b;
    return rc;
  }
  *out = b;
  return DEVQA_OK;
}

int devqa_bank_destroy(devqa_bank* b) {
  if (!b) return DEVQA_OK;
  DEVQA_ON_DEVICE(b->device);
  if (b->data || b->vmm) cudaDeviceSynchronize();
  if (b->shadow_vmm) vmm_destroy(b->shadow_vmm); else if (b->shadow) cudaFree(b->shadow);
  if (b->shadow_scale) cudaFree(b->shadow_scale);
  if (b->vmm) { vmm_destroy(b->vmm); if (b->sumsq) cudaFree(b->sumsq); }
  else if (b->data) { cudaFree(b->data); cudaFree(b->sumsq); }
  if (b->max_sumsq) cudaFree(b->max_sumsq);
  delete b;
  return DEVQA_OK;
}

int64_t devqa_bank_size(const devqa_bank* b) { return b ? b->size : DEVQA_E_INVALID; }
int64_t devqa_bank_capacity(const devqa_bank* b) { return b ? b->capacity : DEVQA_E_INVALID; }
int64_t devqa_bank_dim(const devqa_bank* b) { return b ? b->dim : DEVQA_E_INVALID; }
int devqa_bank_dtype(const devqa_bank* b) { return b ? b->dtype : DEVQA_E_INVALID; }
int devqa_bank_device(const devqa_bank* b) { return b ? b->device : DEVQA_E_INVALID; }
const void* devqa_bank_data(const devqa_bank* b) { return b ? b->data : nullptr; }

int devqa_bank_reserve(devqa_bank* b, int64_t capacity_rows, void* stream) {
  DEVQA_REQUIRE(b != nullptr, "bank is null");
  DEVQA_ON_DEVICE(b->device);
  return bank_grow(b, capacity_rows, static_cast<cudaStream_t>(stream));
}

int devqa_bank_append(devqa_bank* b, const float* rows, int64_t n, int64_t ld, int rows_is_host, void* stream_) {
  DEVQA_REQUIRE(b != nullptr, "bank is null");
  DEVQA_REQUIRE(n >= 0, "negative row count");
  if (n == 0) return DEVQA_OK;
  DEVQA_REQUIRE(rows != nullptr && ld >= b->dim, "bad rows/ld");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DEVQA_ON_DEVICE(b->device);
  int rc = bank_grow(b, b->size + n, stream);
  if (rc) return rc;
  char* dst = static_cast<char*>(b->data) + size_t(b->size) * b->ld * b->esz();
  if (rows_is_host) {
    if (b->dtype == DEVQA_F32) {
      DEVQA_CUDA(cudaMemcpy2DAsync(dst, size_t(b->ld) * 4, rows, size_t(ld) * 4, size_t(b->dim) * 4, size_t(n),
                                   cudaMemcpyHostToDevice, stream));
      DEVQA_CUDA(cudaStreamSynchronize(stream));
    } else {
      float* tmp = nullptr;
      cudaError_t e = cudaMalloc(&tmp, size_t(n) * b->dim * 4);
      if (e != cudaSuccess) { set_error("cudaMalloc staging failed: %s", cudaGetErrorString(e)); return DEVQA_E_NOMEM; }
      DEVQA_CUDA(cudaMemcpy2DAsync(tmp, size_t(b->dim) * 4, rows, size_t(ld) * 4, size_t(b->dim) * 4, size_t(n),
                                   cudaMemcpyHostToDevice, stream));
      rc = convert_rows_launch(tmp, b->dim, dst, b->dtype, b->ld, n, b->dim, stream);
      DEVQA_CUDA(cudaStreamSynchronize(stream));
      cudaFree(tmp);
      if (rc) return rc;
    }
  } else {
    if (b->dtype == DEVQA_F32) {
      DEVQA_CUDA(cudaMemcpy2DAsync(dst, size_t(b->ld) * 4, rows, size_t(ld) * 4, size_t(b->dim) * 4, size_t(n),
                                   cudaMemcpyDeviceToDevice, stream));
    } else {
      if ((rc = convert_rows_launch(rows, ld, dst, b->dtype, b->ld, n, b->dim, stream))) return rc;
    }
  }
  if ((rc = row_sumsq_launch(dst, b->dtype, n, b->dim, b->ld, b->sumsq + b->size, stream, b->max_sumsq))) return rc;
  if ((rc = shadow_update(b, b->size, n, stream))) return rc;
  b->size += n;
  return DEVQA_OK;
}

int devqa_bank_set_shadow(devqa_bank* b, int enable, void* stream_) {
  DEVQA_REQUIRE(b != nullptr, "bank is null");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DEVQA_ON_DEVICE(b->device);
  if (!enable) {
    if (b->shadow_ld) {
      DEVQA_CUDA(cudaDeviceSynchronize());
      if (b->shadow_vmm) vmm_destroy(b->shadow_vmm); else if (b->shadow) cudaFree(b->shadow);
      if (b->shadow_scale) cudaFree(b->shadow_scale);
      b->shadow = nullptr; b->shadow_vmm = nullptr; b->shadow_scale = nullptr; b->shadow_ld = 0; b->shadow_capacity = 0;
    }
    return DEVQA_OK;
  }
  if (b->dtype != DEVQA_F32) { set_error("a shadow only makes sense for an fp32 bank"); return DEVQA_E_UNSUPPORTED; }
  if (b->shadow_ld) return DEVQA_OK;
  static const bool vmm_on = getenv("DEVQA_VMM") == nullptr || atoi(getenv("DEVQA_VMM")) != 0;
  b->shadow_ld = (b->dim + 7) / 8 * 8;                       // rows stay 16-byte aligned
  b->shadow_vmm = vmm_on ? vmm_create(b->device) : nullptr;
  int rc = shadow_update(b, 0, b->size, stream);
  if (rc) {
    if (b->shadow_vmm) vmm_destroy(b->shadow_vmm); else if (b->shadow) cudaFree(b->shadow);
    if (b->shadow_scale) cudaFree(b->shadow_scale);
    b->shadow = nullptr; b->shadow_vmm = nullptr; b->shadow_scale = nullptr; b->shadow_ld = 0; b->shadow_capacity = 0;
  }
  return rc;
}
int devqa_bank_has_shadow(const devqa_bank* b) { return b && b->shadow_ld ? 1 : 0; }

int devqa_bank_read_rows(const devqa_bank* b, int64_t first_row, int64_t n, float* dst, int64_t dst_ld, void* stream_) {
  DEVQA_REQUIRE(b != nullptr, "bank is null");
  DEVQA_REQUIRE(first_row >= 0 && n >= 0 && first_row + n <= b->size, "row range outside the bank");
  if (n == 0) return DEVQA_OK;
  DEVQA_REQUIRE(dst != nullptr && dst_ld >= b->dim, "bad dst/ld");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DEVQA_ON_DEVICE(b->device);
  const char* src = static_cast<const char*>(b->data) + size_t(first_row) * b->ld * b->esz();
  if (b->dtype == DEVQA_F32) {
    DEVQA_CUDA(cudaMemcpy2DAsync(dst, size_t(dst_ld) * 4, src, size_t(b->ld) * 4, size_t(b->dim) * 4, size_t(n),
                                 cudaMemcpyDeviceToDevice, stream));
    return DEVQA_OK;
  }
  return widen_rows_launch(src, b->ld, dst, dst_ld, n, b->dim, stream);
}

int devqa_bank_truncate(devqa_bank* b, int64_t n_rows) {
  DEVQA_REQUIRE(b != nullptr, "bank is null");
  DEVQA_REQUIRE(n_rows >= 0 && n_rows <= b->size, "truncate to %lld outside [0, %lld]", (long long)n_rows,
                (long long)b->size);
  b->size = n_rows;
  return DEVQA_OK;
}

int devqa_bank_set_id_map(devqa_bank* b, int64_t id_base, int64_t chunk_rows, int rank, int world) {
  DEVQA_REQUIRE(b != nullptr, "bank is null");
  DEVQA_REQUIRE(id_base >= 0 && chunk_rows >= 0 && world >= 1 && rank >= 0 && rank < world, "bad id map");
  b->idmap = IdMap{id_base, chunk_rows, rank, world};
  return DEVQA_OK;
}

static int bank_search_common(devqa_bank* b, const float* queries, int64_t nq, int64_t qld, int top_k, unsigned flags,
                              uint64_t* out_keys, float* out_scores, int64_t* out_ids, cudaStream_t stream) {
  DEVQA_REQUIRE(b != nullptr, "bank is null");
  SearchArgs a = bank_args(b, queries, nq, qld, top_k, flags);
  int rc = validate(a);
  if (rc) return rc;
  DEVQA_REQUIRE(nq == 0 || out_keys || (out_scores && out_ids), "output pointers are null");
  DEVQA_ON_DEVICE(b->device);
  Scratch* s = nullptr;
  if ((rc = scratch_for(b->device, &s))) return rc;
  std::lock_guard<std::mutex> lock(s->mu);
  if ((rc = scratch_enter(s, stream))) return rc;
  if ((rc = run_search(s, a, out_keys, out_scores, out_ids, stream))) return rc;
  return scratch_leave(s, stream);
}

int devqa_bank_search(devqa_bank* b, const float* queries, int64_t nq, int64_t qld, int top_k, unsigned flags,
                      float* out_scores, int64_t* out_ids, void* stream) {
  return bank_search_common(b, queries, nq, qld, top_k, flags, nullptr, out_scores, out_ids,
                            static_cast<cudaStream_t>(stream));
}

int devqa_bank_search_packed(devqa_bank* b, const float* queries, int64_t nq, int64_t qld, int top_k, unsigned flags,
                             uint64_t* out_keys, void* stream) {
  DEVQA_REQUIRE(nq == 0 || out_keys, "out_keys is null");
  return bank_search_common(b, queries, nq, qld, top_k, flags, out_keys, nullptr, nullptr,
                            static_cast<cudaStream_t>(stream));
}

int devqa_bank_search_host(devqa_bank* b, const float* qh, int64_t nq, int64_t qld, int top_k, unsigned flags,
                           float* scores_h, int64_t* ids_h, void* stream_) {
  DEVQA_REQUIRE(b != nullptr, "bank is null");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SearchArgs a = bank_args(b, qh, nq, qld, top_k, flags);
  int rc = validate(a);
  if (rc) return rc;
  if (nq == 0) return DEVQA_OK;
  DEVQA_REQUIRE(scores_h && ids_h, "output pointers are null");
  DEVQA_ON_DEVICE(b->device);
  Scratch* s = nullptr;
  if ((rc = scratch_for(b->device, &s))) return rc;
  std::lock_guard<std::mutex> lock(s->mu);
  if ((rc = scratch_enter(s, stream))) return rc;
  float* dq; float* ds; int64_t* di;
  if ((rc = stage_queries(s, qh, nq, b->dim, qld, top_k, &dq, &ds, &di, stream))) return rc;
  a.queries = dq; a.query_ld = b->dim;
  if ((rc = run_search(s, a, nullptr, ds, di, stream))) return rc;
  if ((rc = unstage_results(s, nq, b->dim, top_k, scores_h, ids_h, stream))) return rc;
  return scratch_leave(s, stream);
}

int devqa_bank_select_above(devqa_bank* b, const float* queries, int64_t nq, int64_t qld, const float* thresholds,
                            float scale, unsigned flags, uint8_t* out_mask, int64_t* out_ids, float* out_scores,
                            int64_t cap, int64_t* out_counts, void* stream_) {
  DEVQA_REQUIRE(b != nullptr, "bank is null");
  DEVQA_REQUIRE(nq >= 0 && cap >= 0 && qld >= b->dim, "bad select shape");
  if (nq == 0) return DEVQA_OK;
  DEVQA_REQUIRE(queries && thresholds && out_counts, "null queries / thresholds / out_counts");
  DEVQA_REQUIRE(cap == 0 || (out_ids && out_scores), "out_ids / out_scores are null");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DEVQA_ON_DEVICE(b->device);
  Scratch* s = nullptr;
  int rc = scratch_for(b->device, &s);
  if (rc) return rc;
  std::lock_guard<std::mutex> lock(s->mu);
  if ((rc = scratch_enter(s, stream))) return rc;
  const int64_t n = b->size;
  uint8_t* mask = out_mask;
  if (!mask && n > 0) {
    if ((rc = grow(reinterpret_cast<void**>(&s->selmask), &s->selmask_bytes, size_t(nq) * size_t(n), false))) return rc;
    mask = s->selmask;
  }
  if (n > 0) {
    const int nv_vec = b->dtype == DEVQA_F32 ? 4 : 8;
    const bool vectorised = (b->dim % nv_vec == 0) && (b->ld % nv_vec == 0) && (reinterpret_cast<uintptr_t>(b->data) % 16 == 0);
    const int nv = vectorised ? nv_vec : 1;
    int qt = nq > 1 ? 4 : 1;
    if (scan_stream_smem_bytes(b->dim, nv, qt, 0) > 200 * 1024) qt = 1;
    if (scan_stream_smem_bytes(b->dim, nv, qt, 0) > 220 * 1024) {
      set_error("dim %lld does not fit the scan kernel's shared memory", (long long)b->dim);
      return DEVQA_E_UNSUPPORTED;
    }
    const void* bank_ptr = b->data;
    if (n < 8) {                                           // whole row groups only: zero-padded copy (see run_search)
      const size_t row_bytes = size_t(b->ld) * b->esz();
      if ((rc = grow(&s->tiny, &s->tiny_bytes, 8 * row_bytes, false))) return rc;
      DEVQA_CUDA(cudaMemsetAsync(s->tiny, 0, 8 * row_bytes, stream));
      DEVQA_CUDA(cudaMemcpyAsync(s->tiny, b->data, size_t(n - 1) * row_bytes + size_t(b->dim) * b->esz(),
                                 cudaMemcpyDeviceToDevice, stream));
      bank_ptr = s->tiny;
    }
    const int R = scan_stream_rows_per_group(qt);
    const int64_t groups = (n + R - 1) / R;
    const int64_t want_ctas = (groups + (kScanThreads / 32) - 1) / (kScanThreads / 32);
    const int grid = int(want_ctas < s->sm_count ? (want_ctas < 1 ? 1 : want_ctas) : s->sm_count);
    ScanParams p{};
    p.bank = bank_ptr; p.n_rows = n; p.dim = b->dim; p.bank_ld = b->ld;
    p.queries = queries; p.query_ld = qld; p.n_queries = int32_t(nq);
    p.top_k = 1; p.kcap = 32; p.flags = flags; p.idmap = b->idmap;
    p.sel_thr = thresholds; p.sel_scale = scale; p.sel_mask = mask; p.sel_mask_ld = n;
    for (int64_t q0 = 0; q0 < nq; q0 += qt) {
      p.q0 = int32_t(q0);
      p.nq_pass = int32_t(nq - q0 < qt ? nq - q0 : qt);
      if ((rc = scan_stream_launch(p, b->dtype, vectorised, qt, grid, stream))) return rc;
    }
  }
  SelectParams sp{};
  sp.bank = b->data; sp.bank_dtype = b->dtype; sp.n_rows = n; sp.dim = b->dim; sp.bank_ld = b->ld;
  sp.queries = queries; sp.query_ld = qld; sp.n_queries = int32_t(nq);
  sp.flags = flags; sp.scale = scale; sp.idmap = b->idmap;
  sp.mask = mask; sp.mask_ld = n; sp.out_ids = out_ids; sp.out_scores = out_scores; sp.cap = cap; sp.out_counts = out_counts;
  if ((rc = select_compact_launch(sp, stream))) return rc;
  return scratch_leave(s, stream);
}

// gate (+ optional memory check) on an entered scratch; mem_scores: device [nq] or nullptr
static int gate_common(Scratch* s, const float* queries, int64_t nq, int64_t qld, int64_t dim, const float* w_rand,
                       int64_t w_ld, int64_t m, const float* wo, int64_t wo_ld, int n_classes, unsigned flags,
                       const float* mem_scores, float mem_threshold, float* out_logits, int32_t* out_preds,
                       float* partial, cudaStream_t stream) {
  GateParams g{};
  g.queries = queries; g.query_ld = qld; g.n_queries = int32_t(nq); g.dim = int32_t(dim);
  g.w_rand = w_rand; g.w_ld = w_ld; g.m = int32_t(m); g.wo = wo; g.wo_ld = wo_ld; g.n_classes = n_classes;
  g.flags = flags; g.partial = partial; g.ticket = s->gate_ticket; g.logits = out_logits; g.preds = out_preds;
  g.mem_scores = mem_scores; g.mem_stride = 1; g.mem_threshold = mem_threshold;
  return scope_gate_launch(g, stream);
}

static int gate_validate(const float* queries, int64_t nq, int64_t qld, int64_t dim, const float* w_rand, int64_t w_ld,
                         int64_t m, const float* wo, int64_t wo_ld, int n_classes, const float* out_logits,
                         const int32_t* out_preds) {
  DEVQA_REQUIRE(nq >= 0 && dim > 0 && qld >= dim, "scope gate: bad query shape");
  DEVQA_REQUIRE(m > 0 && m <= (1 << 24) && w_ld >= m && wo_ld >= m, "scope gate: bad W_rand / Wo shape");
  DEVQA_REQUIRE(n_classes >= 1 && n_classes <= 64, "scope gate: n_classes %d unsupported", n_classes);
  DEVQA_REQUIRE(nq == 0 || (queries && w_rand && wo && out_logits && out_preds), "scope gate: null pointers");
  return DEVQA_OK;
}

int devqa_scope_gate(const float* queries, int64_t nq, int64_t qld, int64_t dim, const float* w_rand, int64_t w_ld,
                     int64_t m, const float* wo, int64_t wo_ld, int n_classes, unsigned flags, const float* mem_scores,
                     float mem_threshold, float* out_logits, int32_t* out_preds, void* stream_) {
  int rc = gate_validate(queries, nq, qld, dim, w_rand, w_ld, m, wo, wo_ld, n_classes, out_logits, out_preds);
  if (rc) return rc;
  if (nq == 0) return DEVQA_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int device = 0;
  DEVQA_CUDA(cudaGetDevice(&device));
  Scratch* s = nullptr;
  if ((rc = scratch_for(device, &s))) return rc;
  std::lock_guard<std::mutex> lock(s->mu);
  if ((rc = scratch_enter(s, stream))) return rc;
  GateParams shape{}; shape.m = int32_t(m); shape.n_queries = int32_t(nq); shape.n_classes = n_classes;
  if ((rc = grow(reinterpret_cast<void**>(&s->gate), &s->gate_bytes, scope_gate_partial_floats(shape) * 4, false))) return rc;
  if ((rc = gate_common(s, queries, nq, qld, dim, w_rand, w_ld, m, wo, wo_ld, n_classes, flags, mem_scores, mem_threshold,
                        out_logits, out_preds, s->gate, stream)))
    return rc;
  return scratch_leave(s, stream);
}

int devqa_bank_hice_probe(devqa_bank* b, devqa_bank* memory, const float* queries, const float* mem_queries, int64_t nq,
                          int64_t qld, const float* w_rand, int64_t w_ld, int64_t m, const float* wo, int64_t wo_ld,
                          int n_classes, float mem_threshold, int top_k, unsigned flags, float* out_logits,
                          int32_t* out_preds, float* out_mem_sim, float* out_scores, int64_t* out_ids, void* stream_) {
  DEVQA_REQUIRE(b != nullptr, "bank is null");
  int rc = gate_validate(queries, nq, qld, b->dim, w_rand, w_ld, m, wo, wo_ld, n_classes, out_logits, out_preds);
  if (rc) return rc;
  DEVQA_REQUIRE(!memory || (memory->device == b->device && memory->dim == b->dim), "memory bank: other device or width");
  if (nq == 0) return DEVQA_OK;
  DEVQA_REQUIRE(out_scores && out_ids, "output pointers are null");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SearchArgs a = bank_args(b, queries, nq, qld, top_k, flags);
  if ((rc = validate(a))) return rc;
  DEVQA_ON_DEVICE(b->device);
  Scratch* s = nullptr;
  if ((rc = scratch_for(b->device, &s))) return rc;
  std::lock_guard<std::mutex> lock(s->mu);
  if ((rc = scratch_enter(s, stream))) return rc;
  GateParams shape{}; shape.m = int32_t(m); shape.n_queries = int32_t(nq); shape.n_classes = n_classes;
  const size_t part = scope_gate_partial_floats(shape);
  // scratch: [partial logits | best memory similarity per probe | its row id (int64, 8-byte aligned)]
  const size_t mem_off = (part + 1) / 2 * 2;
  if ((rc = grow(reinterpret_cast<void**>(&s->gate), &s->gate_bytes, (mem_off + size_t(nq) * 4) * 4, false))) return rc;
  float* mem_sim = nullptr;
  if (memory && memory->size > 0) {
    // HICE.py:243-248: probe/||probe|| @ (memory/||memory||)^T, max over the memory = a fused no-eps cosine top-1
    mem_sim = out_mem_sim ? out_mem_sim : s->gate + mem_off;
    int64_t* mem_ids = reinterpret_cast<int64_t*>(s->gate + mem_off + size_t(nq) + (size_t(nq) & 1));
    SearchArgs am = bank_args(memory, mem_queries ? mem_queries : queries, nq, qld, 1,
                              DEVQA_NORMALIZE_BANK | DEVQA_NORMALIZE_QUERY | DEVQA_NOEPS);
    if ((rc = validate(am))) return rc;
    if ((rc = run_search(s, am, nullptr, mem_sim, mem_ids, stream))) return rc;
  } else if (out_mem_sim) {
    DEVQA_CUDA(cudaMemsetAsync(out_mem_sim, 0xff, size_t(nq) * 4, stream));   // NaN: "no memory"
  }
  if ((rc = gate_common(s, queries, nq, qld, b->dim, w_rand, w_ld, m, wo, wo_ld, n_classes,
                        DEVQA_NORMALIZE_QUERY | DEVQA_NOEPS, mem_sim, mem_threshold, out_logits, out_preds, s->gate, stream)))
    return rc;
  if ((rc = run_search(s, a, nullptr, out_scores, out_ids, stream))) return rc;
  return scratch_leave(s, stream);
}

int devqa_topk_merge(const uint64_t* gathered, int n_lists, int64_t nq, int top_k, float* out_scores,
                     int64_t* out_ids, void* stream) {
  DEVQA_REQUIRE(n_lists >= 1 && nq >= 0 && top_k >= 1, "bad merge shape");
  if (top_k > DEVQA_MAX_TOPK) { set_error("devqa_topk_merge: top_k %d exceeds DEVQA_MAX_TOPK", top_k); return DEVQA_E_UNSUPPORTED; }
  DEVQA_REQUIRE(nq == 0 || (gathered && out_scores && out_ids), "null pointers");
  return merge_lists_launch(gathered, n_lists, nq, top_k, top_k, top_k, nullptr, out_scores, out_ids,
                            static_cast<cudaStream_t>(stream));
}

// ---- distributed ----------------------------------------------------------------------------------------
int devqa_dist_unique_id(void* out) {
  DEVQA_REQUIRE(out != nullptr, "out is null");
  const NcclApi* api = nccl_api();
  if (!api) return DEVQA_E_NCCL;
  ncclUniqueId_ id;
  int r = api->GetUniqueId(&id);
  if (r != 0) { set_error("ncclGetUniqueId failed: %s", api->GetErrorString(r)); return DEVQA_E_NCCL; }
  memcpy(out, &id, sizeof(id));
  return DEVQA_OK;
}

int devqa_dist_init(devqa_comm** out, const void* id_bytes, int rank, int world, int device) {
  DEVQA_REQUIRE(out && id_bytes, "null argument");
  DEVQA_REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad rank/world %d/%d", rank, world);
  const NcclApi* api = nccl_api();
  if (!api) return DEVQA_E_NCCL;
  Scratch* s = nullptr;
  int rc = scratch_for(device, &s);
  if (rc) return rc;
  DEVQA_CUDA(cudaSetDevice(device));
  ncclUniqueId_ id;
  memcpy(&id, id_bytes, sizeof(id));
  void* comm = nullptr;
  int r = api->CommInitRank(&comm, world, id, rank);
  if (r != 0) { set_error("ncclCommInitRank failed: %s", api->GetErrorString(r)); return DEVQA_E_NCCL; }
  devqa_comm* c = new (std::nothrow) devqa_comm();
  if (!c) { set_error("host allocation failed"); return DEVQA_E_NOMEM; }
  c->rank = rank; c->world = world; c->device = device; c->nccl = comm;
  *out = c;
  return DEVQA_OK;
}

int devqa_dist_destroy(devqa_comm* c) {
  if (!c) return DEVQA_OK;
  const NcclApi* api = nccl_api();
  if (c->mbox) {
    DeviceGuardLite guard(c->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < c->world; ++r)
      if (r != c->rank && c->peer_mbox[r]) cudaIpcCloseMemHandle(c->peer_mbox[r]);
    cudaFree(c->mbox);
  }
  if (api && c->nccl) api->CommDestroy(c->nccl);
  delete c;
  return DEVQA_OK;
}

// ---- peer-to-peer mailboxes: export -> (handles travel over any host channel) -> connect -------------------------
int devqa_dist_p2p_export(devqa_comm* c, void* out_handle_64_bytes) {
  DEVQA_REQUIRE(c && out_handle_64_bytes, "null argument");
  DEVQA_REQUIRE(c->world <= kP2pMaxWorld, "world %d exceeds the mailbox layout", c->world);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  DEVQA_ON_DEVICE(c->device);
  if (!c->mbox) {
    DEVQA_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->mbox), c->mbox_bytes()));
    DEVQA_CUDA(cudaMemset(c->mbox, 0, c->data_off()));      // flags start at 0; sequence numbers start at 1
    DEVQA_CUDA(cudaDeviceSynchronize());
  }
  cudaIpcMemHandle_t h;
  DEVQA_CUDA(cudaIpcGetMemHandle(&h, c->mbox));
  memcpy(out_handle_64_bytes, &h, 64);
  return DEVQA_OK;
}

int devqa_dist_p2p_connect(devqa_comm* c, const void* handles_world_x_64_bytes) {
  DEVQA_REQUIRE(c && handles_world_x_64_bytes, "null argument");
  DEVQA_REQUIRE(c->mbox != nullptr, "call devqa_dist_p2p_export first");
  DEVQA_ON_DEVICE(c->device);
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) { c->peer_mbox[r] = c->mbox; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const char*>(handles_world_x_64_bytes) + size_t(r) * 64, 64);
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      for (int k = 0; k < r; ++k)
        if (k != c->rank && c->peer_mbox[k]) { cudaIpcCloseMemHandle(c->peer_mbox[k]); c->peer_mbox[k] = nullptr; }
      set_error("cudaIpcOpenMemHandle for rank %d failed: %s (staying on ncclAllGather)", r, cudaGetErrorString(e));
      return DEVQA_E_UNSUPPORTED;
    }
    c->peer_mbox[r] = static_cast<unsigned char*>(ptr);
  }
  c->p2p = true;
  return DEVQA_OK;
}

int devqa_dist_p2p_set(devqa_comm* c, int enable) {
  DEVQA_REQUIRE(c != nullptr, "comm is null");
  DEVQA_REQUIRE(!enable || c->peer_mbox[c->rank] != nullptr, "mailboxes are not connected");
  c->p2p = enable != 0;
  return DEVQA_OK;
}
int devqa_dist_p2p_enabled(const devqa_comm* c) { return c && c->p2p ? 1 : 0; }

int devqa_dist_p2p_timeouts(devqa_comm* c, int64_t* count) {
  DEVQA_REQUIRE(c && count, "null argument");
  *count = 0;
  if (!c->mbox) return DEVQA_OK;
  DEVQA_ON_DEVICE(c->device);
  uint32_t v = 0;
  DEVQA_CUDA(cudaDeviceSynchronize());
  DEVQA_CUDA(cudaMemcpy(&v, c->mbox, 4, cudaMemcpyDeviceToHost));
  *count = v;
  return DEVQA_OK;
}
int devqa_dist_rank(const devqa_comm* c) { return c ? c->rank : DEVQA_E_INVALID; }
int devqa_dist_world(const devqa_comm* c) { return c ? c->world : DEVQA_E_INVALID; }

static int dist_search(devqa_bank* b, devqa_comm* c, SearchArgs a, Scratch* s, float* out_scores, int64_t* out_ids,
                       cudaStream_t stream) {
  if (a.top_k > DEVQA_MAX_TOPK) {
    set_error("sharded search: top_k %d exceeds DEVQA_MAX_TOPK (%d)", a.top_k, DEVQA_MAX_TOPK);
    return DEVQA_E_UNSUPPORTED;
  }
  const NcclApi* api = nccl_api();
  if (!api) return DEVQA_E_NCCL;
  const size_t per_rank = size_t(a.n_queries) * a.top_k;
  int rc = grow(reinterpret_cast<void**>(&s->keys), &s->keys_bytes, per_rank * 8, false);
  if (rc) return rc;
  if (c->p2p && a.n_queries <= kP2pMaxQ && int64_t(per_rank) <= kP2pMaxKeys) {
    // small batches (the latency-bound regime): each rank's merge kernel pushes its keys into the peers' mailboxes
    // and waits for theirs -- no collective launch (the decision depends on the call's shape only: every rank takes
    // the same branch, and the sequence numbers stay in lockstep)
    if ((rc = run_search(s, a, s->keys, nullptr, nullptr, stream))) return rc;
    P2pExchange x{};
    x.local_keys = s->keys; x.rank = c->rank; x.world = c->world;
    x.header = reinterpret_cast<uint32_t*>(c->mbox);
    for (int r = 0; r < c->world; ++r) {
      x.peer_flags[r] = reinterpret_cast<uint32_t*>(c->peer_mbox[r] + c->flags_off());
      x.peer_data[r] = reinterpret_cast<uint64_t*>(c->peer_mbox[r] + c->data_off());
    }
    return merge_lists_launch(x.peer_data[c->rank], c->world, a.n_queries, a.top_k, a.top_k, a.top_k, nullptr, out_scores,
                              out_ids, stream, nullptr, nullptr, &x);
  }
  if ((rc = grow(reinterpret_cast<void**>(&s->gathered), &s->gathered_bytes, per_rank * 8 * c->world, false))) return rc;
  if ((rc = run_search(s, a, s->keys, nullptr, nullptr, stream))) return rc;
  int r = api->AllGather(s->keys, s->gathered, per_rank, /*ncclUint64*/ 5, c->nccl, stream);
  if (r != 0) { set_error("ncclAllGather failed: %s", api->GetErrorString(r)); return DEVQA_E_NCCL; }
  return merge_lists_launch(s->gathered, c->world, a.n_queries, a.top_k, a.top_k, a.top_k, nullptr, out_scores,
                            out_ids, stream);
}

int devqa_bank_search_dist(devqa_bank* b, devqa_comm* c, const float* queries, int64_t nq, int64_t qld, int top_k,
                           unsigned flags, float* out_scores, int64_t* out_ids, void* stream_) {
  DEVQA_REQUIRE(b && c, "null bank/comm");
  DEVQA_REQUIRE(b->device == c->device, "bank and communicator live on different devices");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SearchArgs a = bank_args(b, queries, nq, qld, top_k, flags);
  int rc = validate(a);
  if (rc) return rc;
  if (nq == 0) return DEVQA_OK;
  DEVQA_REQUIRE(out_scores && out_ids, "output pointers are null");
  DEVQA_ON_DEVICE(b->device);
  Scratch* s = nullptr;
  if ((rc = scratch_for(b->device, &s))) return rc;
  std::lock_guard<std::mutex> lock(s->mu);
  if ((rc = scratch_enter(s, stream))) return rc;
  if ((rc = dist_search(b, c, a, s, out_scores, out_ids, stream))) return rc;
  return scratch_leave(s, stream);
}

int devqa_bank_search_dist_host(devqa_bank* b, devqa_comm* c, const float* qh, int64_t nq, int64_t qld, int top_k,
                                unsigned flags, float* scores_h, int64_t* ids_h, void* stream_) {
  DEVQA_REQUIRE(b && c, "null bank/comm");
  DEVQA_REQUIRE(b->device == c->device, "bank and communicator live on different devices");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SearchArgs a = bank_args(b, qh, nq, qld, top_k, flags);
  int rc = validate(a);
  if (rc) return rc;
  if (nq == 0) return DEVQA_OK;
  DEVQA_REQUIRE(scores_h && ids_h, "output pointers are null");
  DEVQA_ON_DEVICE(b->device);
  Scratch* s = nullptr;
  if ((rc = scratch_for(b->device, &s))) return rc;
  std::lock_guard<std::mutex> lock(s->mu);
  if ((rc = scratch_enter(s, stream))) return rc;
  float* dq; float* ds; int64_t* di;
  if ((rc = stage_queries(s, qh, nq, b->dim, qld, top_k, &dq, &ds, &di, stream))) return rc;
  a.queries = dq; a.query_ld = b->dim;
  if ((rc = dist_search(b, c, a, s, ds, di, stream))) return rc;
  if ((rc = unstage_results(s, nq, b->dim, top_k, scores_h, ids_h, stream))) return rc;
  return scratch_leave(s, stream);
}

// ---- synthetic ----------------------------------------------------------------------------------------------
int devqa_fill_synthetic(float* dst, int64_t n_rows, int64_t dim, int64_t ld, uint64_t seed, int64_t first_global_row,
                         int kind, void* stream) {
  DEVQA_REQUIRE(dst && n_rows >= 0 && dim > 0 && ld >= dim, "bad fill arguments");
  DEVQA_REQUIRE(kind == 0 || kind == 1, "unknown synthetic kind %d", kind);
  return fill_synthetic_launch(dst, DEVQA_F32, n_rows, dim, ld, seed, first_global_row, kind,
                               static_cast<cudaStream_t>(stream));
}

int devqa_bank_append_synthetic(devqa_bank* b, int64_t n_rows, uint64_t seed, int64_t first_global_row, int kind,
                                void* stream_) {
  DEVQA_REQUIRE(b != nullptr && n_rows >= 0, "bad arguments");
  DEVQA_REQUIRE(kind == 0 || kind == 1, "unknown synthetic kind %d", kind);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DEVQA_ON_DEVICE(b->device);
  int rc = bank_grow(b, b->size + n_rows, stream);
  if (rc) return rc;
  char* dst = static_cast<char*>(b->data) + size_t(b->size) * b->ld * b->esz();
  if ((rc = fill_synthetic_launch(dst, b->dtype, n_rows, b->dim, b->ld, seed, first_global_row, kind, stream))) return rc;
  if ((rc = row_sumsq_launch(dst, b->dtype, n_rows, b->dim, b->ld, b->sumsq + b->size, stream, b->max_sumsq))) return rc;
  if ((rc = shadow_update(b, b->size, n_rows, stream))) return rc;
  b->size += n_rows;
  return DEVQA_OK;
}

}  // extern "C"
