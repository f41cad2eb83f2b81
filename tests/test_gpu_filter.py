"""The TF32 filter + exact re-scoring stage of the tcgen05 scan (fp32 banks, batches > 4 queries):
its error bound, its bit-identity with the CUDA-core stream kernel, and the exact fallback for queries
whose candidate list is not provably complete.  GPU only."""
import ctypes as C

import numpy as np
import pytest
import torch

import devqa_b200
from devqa_b200 import FactBank
from oracle import semantic_search as ss

pytestmark = pytest.mark.gpu
L = devqa_b200._lib
GAP, RTOL, ATOL = 1e-5, 1e-5, 1e-6
FILTER = L.FORCE_TENSOR | L.FORCE_FILTER     # small banks would otherwise be scanned with 3xTF32 alone


def _randn(n, d, seed, unit=True):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, d, generator=g)
    return torch.nn.functional.normalize(x, dim=1) if unit else x


def _hidden_like(n, d, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, d, generator=g) * torch.exp(0.5 * torch.randn(n, 1, generator=g))
    x[:, [1, d // 3, d // 2, d - 2]] *= 50.0
    return x


def _fallbacks(device=0):
    n = C.c_int64()
    L.check(L.lib().devqa_filter_fallbacks(device, C.byref(n)))
    return n.value


def _record(name, value):
    """Keep measured error / bound pairs next to the test log (gpurun_out/ travels back from the GPU box)."""
    import json, os
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        path = os.path.join(out, "filter_error_bound.json")
        data = json.load(open(path)) if os.path.exists(path) else {}
        data[name] = value
        json.dump(data, open(path, "w"), indent=1)


def _violations(device=0):
    n = C.c_int64()
    L.check(L.lib().devqa_filter_bound_violations(device, C.byref(n)))
    return n.value


def relerr(d, bf16=False):
    """csrc/scan.cuh tc_filter_relerr, as the library itself reports it"""
    return float(L.lib().devqa_filter_relerr(int(d), L.BF16 if bf16 else L.F32))


def test_relerr_is_what_the_header_documents():
    assert relerr(768) == pytest.approx(1.5 * 2.0 ** -10 + 2.0 ** -19 + 768 * 2.0 ** -22, rel=1e-6)
    # bf16 has 8 significand bits: round-to-nearest is off by up to 2^-8 relative (not 2^-9)
    assert relerr(768, bf16=True) == pytest.approx(2.0 ** -8 + 2.0 ** -19 + 768 * 2.0 ** -22, rel=1e-6)
    x = torch.tensor([1.0 + 2.0 ** -8 + 2.0 ** -12])
    assert abs(float(x.to(torch.bfloat16).float() - x) / float(x)) > 1.01 * 2.0 ** -9


@pytest.mark.parametrize("n,d,nq,k,nb,nq_mode", [
    (5000, 768, 64, 32, "eps", "eps"), (5000, 768, 128, 32, "eps", "eps"), (3000, 384, 20, 16, "eps", "eps"),
    (4097, 100, 9, 5, None, None), (2000, 4096, 40, 4, "eps", "eps"), (300, 36, 7, 8, "eps", None),
    (6000, 1000, 130, 64, "noeps", "noeps"), (129, 64, 200, 3, None, "eps")])
def test_filter_matches_stream_kernel_bit_for_bit(n, d, nq, k, nb, nq_mode):
    """Re-scoring repeats scan_stream's arithmetic, so the kernel choice is invisible in results."""
    bank, q = _randn(n, d, n + d, unit=False), _randn(nq, d, nq + k, unit=False)
    fb = FactBank.from_embeddings(bank)
    _fallbacks()
    a = fb.search(q.cuda(), k, normalize_bank=nb, normalize_query=nq_mode, flags=FILTER)
    assert _fallbacks() == 0
    b = fb.search(q.cuda(), k, normalize_bank=nb, normalize_query=nq_mode, flags=L.FORCE_STREAM)
    assert torch.equal(a[1], b[1])
    assert torch.equal(a[0], b[0])
    truth = ss.truth_scores_fp64(q, bank, normalize_bank=bool(nb), normalize_query=bool(nq_mode))
    kk = min(k, n)
    ss.compare_topk(a[0].cpu().numpy()[:, :kk], a[1].cpu().numpy()[:, :kk], truth, k, gap=GAP, rtol=RTOL, atol=ATOL)
    fb.close()


def test_filter_on_caller_owned_tensor_without_cached_norms():
    """devqa_search_raw has no cached ||row||^2: the norm warps supply them (and max ||row|| for dot scores)."""
    bank, q = _hidden_like(3000, 512, 7), _hidden_like(33, 512, 8)
    dev_bank, dev_q = bank.cuda(), q.cuda()
    for normalise in (True, False):
        score_fn = devqa_b200.util.cos_sim if normalise else devqa_b200.util.dot_score
        s, i = devqa_b200.util.semantic_search_arrays(dev_q, dev_bank, top_k=10, score_function=score_fn, flags=FILTER)
        truth = ss.truth_scores_fp64(q, bank, normalize_bank=normalise, normalize_query=normalise)
        ss.compare_topk(s.cpu().numpy(), i.cpu().numpy(), truth, 10, gap=GAP, rtol=RTOL, atol=ATOL)
        s2, i2 = devqa_b200.util.semantic_search_arrays(dev_q, dev_bank, top_k=10, score_function=score_fn, flags=L.FORCE_STREAM)
        if normalise:                                        # the filter ran (cos_sim): bit-identical to the stream kernel
            assert torch.equal(i, i2) and torch.equal(s, s2)


@pytest.mark.parametrize("kind", ["randn", "positive", "hidden", "hidden4096"])
def test_filter_error_bound_holds(kind):
    """The filter ranks rows by u = approximate score + E and everything rests on  exact <= u <= exact + 2E,
    E = relerr(D) * ||row|| * ||query||.  Checked on adversarial inputs: all-positive vectors make every TF32
    truncation error point the same way; hidden-state rows have x50 outlier columns."""
    g = torch.Generator().manual_seed(99)
    if kind == "randn":
        bank, q = _randn(4000, 768, 1, unit=False), _randn(48, 768, 2, unit=False)
    elif kind == "positive":
        bank, q = torch.rand(4000, 768, generator=g) + 0.5, torch.rand(48, 768, generator=g) + 0.5
    elif kind == "hidden":
        bank, q = _hidden_like(4000, 768, 3), _hidden_like(48, 768, 4)
    else:
        bank, q = _hidden_like(2000, 4096, 5), _hidden_like(32, 4096, 6)
    d = bank.shape[1]
    fb = FactBank.from_embeddings(bank)
    for nb, nqm in (("eps", "eps"), (None, None)):
        u, i = fb.search(q.cuda(), 32, normalize_bank=nb, normalize_query=nqm, flags=FILTER | L.DEBUG_FILTER_ONLY)
        u, i = u.cpu().double().numpy(), i.cpu().numpy()
        truth = ss.truth_scores_fp64(q, bank, normalize_bank=bool(nb), normalize_query=bool(nqm))
        exact = np.take_along_axis(truth, i, axis=1)
        bn = np.linalg.norm(bank.double().numpy(), axis=1)
        qn = np.linalg.norm(q.double().numpy(), axis=1)
        scale = np.ones_like(exact) if nb else bn[i] * qn[:, None]
        e = relerr(d) * scale
        err = np.abs(u - e - exact) / scale                     # |approximate - exact| per unit of ||row|| ||query||
        _record(f"{kind}_{'cos' if nb else 'dot'}_d{d}", {"max_err": float(err.max()), "bound": relerr(d)})
        assert (u >= exact).all(), (kind, nb, float((exact - u).max()))
        assert (u <= exact + 2.0001 * e).all(), (kind, nb)
        assert err.max() > 1e-6          # these really are approximate scores
    fb.close()


# ---- the bound on constructed worst cases -----------------------------------------------------------------------
def _bits_to_f32(bits):
    return torch.from_numpy(np.asarray(bits, dtype=np.uint32).view(np.float32).copy())


def _worst_case(n, nq, d, seed, bf16):
    """Rows and queries whose per-element rounding errors all point the same way and for which Cauchy-Schwarz is
    tight (query j is parallel to row j): every element is a power of two times a fixed mantissa,
      fp32 bank:  row mantissa  = low 13 bits all ones (TF32 truncation loses 2^-10 relative, the maximum),
                  query mantissa = one below the TF32 rounding midpoint (rounds DOWN by 2^-11, the maximum);
      bf16 bank:  row = the power of two itself (exact in bf16), query mantissa = one below the bf16 midpoint
                  (rounds down by 2^-8).
    Exponents vary per column (8 values), so the tensor core's truncating accumulation has work to do as well."""
    rng = np.random.default_rng(seed)
    e = rng.integers(120, 128, size=(n, d), dtype=np.uint32)            # 2^-7 .. 2^0
    row_m = np.uint32(0) if bf16 else np.uint32(0x1FFF)
    q_m = np.uint32(0x7FFF) if bf16 else np.uint32(0x0FFF)
    rows = _bits_to_f32((e << np.uint32(23)) | row_m)
    eq = e[rng.integers(0, n, size=nq)] if nq > n else e[:nq]
    queries = _bits_to_f32((eq << np.uint32(23)) | q_m)
    return rows, queries


@pytest.mark.parametrize("d", [768, 4096, 16384])
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_filter_error_bound_on_the_constructed_worst_case(d, dtype):
    """positive rows with low-13-bits-all-ones mantissas x positive queries just below the TF32
    midpoints push the per-element error to 2^-10 + 2^-11 with Cauchy-Schwarz tight; what is left of the bound has to
    cover the tensor core's accumulation.  Measured error / bound is recorded per case (gpurun_out/filter_error_bound.json)."""
    bf = dtype == "bf16"
    n, nq, k = 64, 32, 64                               # k = n: the bound of EVERY pair comes back, the parallel ones included
    bank, q = _worst_case(n, nq, d, 1000 + d, bf)
    fb = FactBank.from_embeddings(bank, dtype=dtype)
    b64, q64 = bank.double().numpy(), q.double().numpy()
    bn, qn = np.linalg.norm(b64, axis=1), np.linalg.norm(q64, axis=1)
    for nb, nqm in ((None, None), ("eps", None)):
        u, i = fb.search(q.cuda(), k, normalize_bank=nb, normalize_query=nqm, flags=FILTER | L.DEBUG_FILTER_ONLY)
        u, i = u.cpu().double().numpy(), i.cpu().numpy()
        exact = np.take_along_axis(q64 @ b64.T / (bn[None, :] if nb else 1.0), i, axis=1)
        scale = qn[:, None] * (np.ones_like(exact) if nb else bn[i])
        e = relerr(d, bf) * scale
        err = np.abs(u - e - exact) / scale
        inputs_only = (2.0 ** -8) if bf else (2.0 ** -10 + 2.0 ** -11)
        _record(f"worstcase_{dtype}_{'cos' if nb else 'dot'}_d{d}",
                {"max_err": float(err.max()), "bound": relerr(d, bf), "input_rounding_alone": inputs_only})
        assert (u >= exact).all(), (d, dtype, nb, float(((exact - u) / scale).max()))
        assert (u <= exact + 2.0001 * e).all(), (d, dtype, nb)
        assert err.max() > 0.9 * inputs_only, "the construction is supposed to come close to the input-rounding bound"
        assert err.max() < relerr(d, bf)
    # and the real thing: filter + exact re-scoring equals the stream kernel, canary silent
    _fallbacks(); _violations()
    a = fb.search(q.cuda(), 32, normalize_bank=None, normalize_query=None, flags=FILTER)
    assert _violations() == 0
    _fallbacks()
    b = fb.search(q.cuda(), 32, normalize_bank=None, normalize_query=None, flags=L.FORCE_STREAM)
    # (row j and query j are parallel, scores of other pairs are close: the fallback may run; results must agree in ids
    # wherever the stream kernel's scores differ)
    sa, sb = a[0].cpu().numpy(), b[0].cpu().numpy()
    np.testing.assert_allclose(sa, sb, rtol=1e-5)
    fb.close()


@pytest.mark.parametrize("kind", ["randn", "positive", "hidden", "worst"])
@pytest.mark.parametrize("d", [36, 768, 4096])
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_filter_bound_exhaustively_on_tiny_banks(kind, d, dtype):
    """test_filter_error_bound_holds only sees rows that were RETURNED.  With N = 64 rows and
    k = 64 the filter returns its upper bound for EVERY (row, query) pair, so every pair is held against fp64 and
    against the stream kernel's own fp32 score (the quantity the proof is really about)."""
    bf = dtype == "bf16"
    n, nq = 64, 48
    g = torch.Generator().manual_seed(7 * d + len(kind))
    if kind == "randn":
        bank, q = _randn(n, d, d + 1, unit=False), _randn(nq, d, d + 2, unit=False)
    elif kind == "positive":
        bank, q = torch.rand(n, d, generator=g) + 0.5, torch.rand(nq, d, generator=g) + 0.5
    elif kind == "hidden":
        bank, q = _hidden_like(n, d, d + 3), _hidden_like(nq, d, d + 4)
    else:
        bank, q = _worst_case(n, nq, d, d + 5, bf)
    fb = FactBank.from_embeddings(bank, dtype=dtype)
    stored = bank.to(torch.bfloat16).float() if bf else bank
    b64, q64 = stored.double().numpy(), q.double().numpy()
    bn, qn = np.linalg.norm(b64, axis=1), np.linalg.norm(q64, axis=1)
    worst = 0.0
    for nb, nqm in ((None, None), ("eps", "eps"), ("eps", None)):
        u, i = fb.search(q.cuda(), n, normalize_bank=nb, normalize_query=nqm, flags=FILTER | L.DEBUG_FILTER_ONLY)
        s, j = fb.search(q.cuda(), n, normalize_bank=nb, normalize_query=nqm, flags=L.FORCE_STREAM)
        u, i, s, j = u.cpu().double().numpy(), i.cpu().numpy(), s.cpu().double().numpy(), j.cpu().numpy()
        assert (np.sort(i, axis=1) == np.arange(n)[None, :]).all(), "every row must come back"
        u_by_row = np.empty_like(u); np.put_along_axis(u_by_row, i, u, axis=1)
        s_by_row = np.empty_like(s); np.put_along_axis(s_by_row, j, s, axis=1)
        truth = ss.truth_scores_fp64(q, stored, normalize_bank=bool(nb), normalize_query=bool(nqm))
        qscale = np.ones(nq) if nqm else qn
        scale = qscale[:, None] * (np.ones((nq, n)) if nb else bn[None, :])
        e = relerr(d, bf) * scale
        assert (u_by_row >= truth).all(), (kind, d, dtype, nb, float(((truth - u_by_row) / scale).max()))
        assert (u_by_row >= s_by_row).all(), (kind, d, dtype, nb)
        assert (u_by_row <= truth + 2.0001 * e).all() and (u_by_row <= s_by_row + 2.0001 * e).all(), (kind, d, dtype, nb)
        worst = max(worst, float((np.abs(u_by_row - e - truth) / scale).max()))
    _record(f"exhaustive_{kind}_{dtype}_d{d}", {"max_err": worst, "bound": relerr(d, bf), "pairs": 3 * n * nq})
    fb.close()


def test_canary_catches_a_violated_bound():
    """DEVQA_DEBUG_TIGHT_BOUND runs the filter with 1/64 of its error bound: exact scores then do fall outside
    [u - 2E, u], the re-scoring kernels must notice, count it, and send the query to the exact fallback."""
    g = torch.Generator().manual_seed(5)
    bank, q = torch.rand(30000, 768, generator=g) + 0.5, torch.rand(40, 768, generator=g) + 0.5
    fb = FactBank.from_embeddings(bank)
    for n_rows in (30000, 3000):                       # refine from the per-CTA lists / from the merged list
        fb.truncate(n_rows)
        _fallbacks(); _violations()
        s, i = fb.search(q.cuda(), 16, flags=FILTER | L.DEBUG_TIGHT_BOUND)
        assert _violations() > 0
        assert _fallbacks() > 0
        fb.search(q.cuda(), 16, flags=FILTER)
        assert _violations() == 0
    fb.close()


@pytest.mark.parametrize("kind", ["positive", "hidden", "constant_offset"])
def test_bf16_filter_error_bound_holds(kind):
    """the bf16 filter's bound was 2x too small; adversarial inputs for it are queries whose
    components carry aligned rounding errors (constant / offset hidden states) against all-positive rows."""
    g = torch.Generator().manual_seed(77)
    d = 768
    if kind == "positive":
        bank, q = torch.rand(4000, d, generator=g) + 0.5, torch.rand(48, d, generator=g) + 0.5
    elif kind == "hidden":
        bank, q = _hidden_like(4000, d, 13).abs(), _hidden_like(48, d, 14).abs()
    else:
        bank = torch.rand(4000, d, generator=g) + 0.5
        q = (1.0 + 2.0 ** -8 + 2.0 ** -12) * torch.ones(48, d) * torch.exp2(torch.randint(-3, 1, (48, 1), generator=g).float())
    fb = FactBank.from_embeddings(bank, dtype="bf16")
    stored = bank.to(torch.bfloat16).float()
    for nb, nqm in (("eps", "eps"), (None, None)):
        u, i = fb.search(q.cuda(), 32, normalize_bank=nb, normalize_query=nqm, flags=FILTER | L.DEBUG_FILTER_ONLY)
        u, i = u.cpu().double().numpy(), i.cpu().numpy()
        truth = ss.truth_scores_fp64(q, stored, normalize_bank=bool(nb), normalize_query=bool(nqm))
        exact = np.take_along_axis(truth, i, axis=1)
        bn = np.linalg.norm(stored.double().numpy(), axis=1)
        qn = np.linalg.norm(q.double().numpy(), axis=1)
        scale = np.ones_like(exact) if nb else bn[i] * qn[:, None]
        e = relerr(d, True) * scale
        err = np.abs(u - e - exact) / scale
        _record(f"bf16_{kind}_{'cos' if nb else 'dot'}_d{d}", {"max_err": float(err.max()), "bound": relerr(d, True)})
        assert (u >= exact).all(), (kind, nb, float(((exact - u) / scale).max()))
        assert (u <= exact + 2.0001 * e).all(), (kind, nb)
    _violations()
    a = fb.search(q.cuda(), 32, flags=FILTER)
    assert _violations() == 0
    b = fb.search(q.cuda(), 32, flags=L.FORCE_STREAM)
    assert torch.equal(a[1], b[1]) and torch.equal(a[0], b[0])
    fb.close()


def test_duplicated_rows_take_the_exact_fallback():
    """200 identical consecutive rows.  Persistent list epilogue (DEVQA_DEBUG_NO_ONESHOT): one CTA's list fills up with
    ties, rows it turns away are not provably irrelevant, and the 3xTF32 scan redoes that query.  One-shot mode (the
    default at this size): all 200 ties are emitted as candidates and re-scored, no fallback needed.  Ties resolve to
    ascending ids either way."""
    bank = _randn(5000, 256, 11)
    bank[1000:1200] = bank[1000]
    q = torch.cat([bank[1000:1001] + 0.01 * _randn(1, 256, 12), _randn(15, 256, 13)])
    fb = FactBank.from_embeddings(bank)
    _fallbacks()
    s, i = fb.search(q.cuda(), 32, flags=FILTER | L.DEBUG_NO_ONESHOT)
    n_fb = _fallbacks()
    assert 1 <= n_fb <= 2
    assert i[0].tolist() == list(range(1000, 1032))
    s1, i1 = fb.search(q.cuda(), 32, flags=FILTER)
    assert _fallbacks() == 0 and torch.equal(i1, i) and torch.equal(s1, s)
    truth = ss.truth_scores_fp64(q, bank)
    ss.compare_topk(s.cpu().numpy(), i.cpu().numpy(), truth, 32, gap=GAP, rtol=RTOL, atol=ATOL)
    # every query -- answered by the filter + re-scoring, or by the fallback and then re-scored in the same
    # arithmetic -- is bit-identical to the stream kernel's answer
    s2, i2 = fb.search(q.cuda(), 32, flags=L.FORCE_STREAM)
    assert torch.equal(i, i2) and torch.equal(s, s2)
    fb.close()


@pytest.mark.parametrize("n,d,nq,k", [(40000, 128, 70, 32), (25000, 384, 130, 5), (60000, 64, 300, 16), (19000, 768, 64, 64)])
def test_filter_midsize_banks_refine_from_the_per_cta_lists(n, d, nq, k):
    """>= 148 tiles: the CTAs exchange their bests and refine picks its candidates out of the per-CTA lists."""
    bank, q = _randn(n, d, n + 1, unit=False), _randn(nq, d, nq + 2, unit=False)
    fb = FactBank.from_embeddings(bank)
    _fallbacks()
    a = fb.search(q.cuda(), k, flags=FILTER)
    nfb = _fallbacks()
    b = fb.search(q.cuda(), k, flags=L.FORCE_STREAM)
    truth = ss.truth_scores_fp64(q, bank)
    ss.compare_topk(a[0].cpu().numpy(), a[1].cpu().numpy(), truth, k, gap=GAP, rtol=RTOL, atol=ATOL)
    if nfb == 0:
        assert torch.equal(a[1], b[1]) and torch.equal(a[0], b[0])
    assert nfb <= nq // 8
    fb.close()


def test_duplicated_rows_midsize_take_the_exact_fallback():
    """400 identical rows in a mid-size bank.  List epilogue (DEVQA_DEBUG_NO_EMIT): one CTA's list overflows with ties
    -> flagged -> exact fallback.  Emit mode (default): all 400 ties are candidates and re-scored, no fallback."""
    bank = _randn(30000, 128, 61)
    bank[5000:5400] = bank[5000]
    q = torch.cat([bank[5000:5001] + 0.01 * _randn(1, 128, 62), _randn(31, 128, 63)])
    fb = FactBank.from_embeddings(bank)
    _fallbacks()
    s, i = fb.search(q.cuda(), 32, flags=FILTER | L.DEBUG_NO_EMIT)
    assert _fallbacks() >= 1
    assert i[0].tolist() == list(range(5000, 5032))
    ss.compare_topk(s.cpu().numpy(), i.cpu().numpy(), ss.truth_scores_fp64(q, bank), 32, gap=GAP, rtol=RTOL, atol=ATOL)
    s2, i2 = fb.search(q.cuda(), 32, flags=L.FORCE_STREAM)
    assert torch.equal(i, i2) and torch.equal(s, s2)
    s3, i3 = fb.search(q.cuda(), 32, flags=FILTER)
    assert _fallbacks() == 0 and torch.equal(i3, i2) and torch.equal(s3, s2)
    fb.close()


def test_clustered_scores_stay_exact():
    """All-positive rows: cosine scores crowd into a narrow band, most queries fail the proof."""
    g = torch.Generator().manual_seed(21)
    bank, q = torch.rand(20000, 384, generator=g), torch.rand(24, 384, generator=g)
    fb = FactBank.from_embeddings(bank)
    s, i = fb.search(q.cuda(), 16, flags=FILTER)
    truth = ss.truth_scores_fp64(q, bank)
    ss.compare_topk(s.cpu().numpy(), i.cpu().numpy(), truth, 16, gap=GAP, rtol=RTOL, atol=ATOL)
    fb.close()


def test_nan_rows_and_zero_rows():
    bank = _randn(3000, 128, 31, unit=False)
    bank[17] = 0
    bank[99, 5] = float("nan")
    q = _randn(10, 128, 32, unit=False)
    fb = FactBank.from_embeddings(bank)
    s, i = fb.search(q.cuda(), 4, normalize_bank="noeps", normalize_query="noeps", flags=FILTER)
    assert set(i[:, :2].flatten().tolist()) == {17, 99} and torch.isnan(s[:, :2]).all()
    s, i = fb.search(q.cuda(), 4, flags=FILTER)               # eps clamp: the zero row scores 0, NaN row first
    assert (i[:, 0] == 99).all() and torch.isnan(s[:, 0]).all()
    fb.close()


def test_large_batch_many_passes_and_chunks():
    """Q = 2100: three scratch chunks of 1024 queries, 128-query passes, several launches."""
    bank, q = _randn(20000, 128, 41), _randn(2100, 128, 42)
    fb = FactBank.from_embeddings(bank)
    s, i = fb.search(q.cuda(), 8)
    truth = ss.truth_scores_fp64(q, bank)
    ss.compare_topk(s.cpu().numpy(), i.cpu().numpy(), truth, 8, gap=GAP, rtol=RTOL, atol=ATOL)
    keys = fb.search_packed(q.cuda(), 8)
    s3, i3 = devqa_b200.bank.topk_merge(keys.unsqueeze(0).contiguous(), 8)
    assert torch.equal(i3, i) and torch.equal(s3, s)
    fb.close()


def test_sharded_filter_equals_single_bank():
    """Chunked round-robin id map: re-scoring inverts it to find the local row."""
    bank, q = _randn(9000, 192, 51), _randn(40, 192, 52)
    whole = FactBank.from_embeddings(bank)
    ref = whole.search(q.cuda(), 16, flags=L.FORCE_STREAM)
    chunk, world = 500, 3
    keys = []
    for rank in range(world):
        rows = torch.cat([bank[c * chunk:(c + 1) * chunk] for c in range(rank, 18, world)])
        fb = FactBank.from_embeddings(rows)
        fb.set_id_map(0, chunk, rank, world)
        keys.append(fb.search_packed(q.cuda(), 16, flags=FILTER))
        fb.close()
    s, i = devqa_b200.bank.topk_merge(torch.stack(keys).contiguous(), 16)
    assert torch.equal(i, ref[1]) and torch.equal(s, ref[0])
    whole.close()


def test_full_size_bank_has_no_fallbacks():
    """1M x 768 unit rows, Q = 128, k = 32: the filter's candidate list (K' = 64) is provably complete for
    every query, i.e. the fast path is the one that runs at the benchmark shapes."""
    fb = FactBank(768, capacity=1_000_000)
    fb.append_synthetic(1_000_000, 3234)
    q = _randn(128, 768, 7678).cuda()
    _fallbacks()
    a = fb.search(q, 32)
    assert _fallbacks() == 0
    b = fb.search(q, 32, flags=L.FORCE_TENSOR | L.FORCE_TF32X3)
    assert (a[1] == b[1]).float().mean() > 0.995       # the 3xTF32 scores differ in the last bits: near-ties may swap
    torch.testing.assert_close(a[0], b[0], rtol=1e-5, atol=1e-6)
    c = fb.search(q[:8], 32, flags=L.FORCE_STREAM)
    assert torch.equal(a[1][:8], c[1]) and torch.equal(a[0][:8], c[0])
    fb.close()


@pytest.mark.parametrize("seed", range(24))
def test_filter_random_shapes_against_the_stream_kernel(seed):
    """Randomised shapes / modes: the tensor-core filter path returns what the CUDA-core stream kernel returns, bit
    for bit (queries that took the 3xTF32 fallback: within the fp32 contract)."""
    import random
    rnd = random.Random(1000 + seed)
    n = rnd.choice([1, 7, 127, 129, 1000, 4097, 20000, 50001])
    d = 4 * rnd.randint(1, 256)
    nq = rnd.choice([5, 16, 17, 63, 64, 65, 130, 257, 300])
    k = rnd.choice([1, 2, 5, 16, 31, 32, 33, 64])
    nb, nqm = rnd.choice([("eps", "eps"), ("noeps", "noeps"), (None, None), ("eps", None), (None, "eps")])
    unit = rnd.random() < 0.5
    bank, q = _randn(n, d, 7 * seed + 1, unit=unit), _randn(nq, d, 7 * seed + 2, unit=unit)
    if not unit and rnd.random() < 0.5:
        bank = bank * torch.exp(torch.randn(n, 1, generator=torch.Generator().manual_seed(seed)))
    fb = FactBank.from_embeddings(bank)
    _fallbacks()
    a = fb.search(q.cuda(), k, normalize_bank=nb, normalize_query=nqm, flags=FILTER)
    nfb = _fallbacks()
    b = fb.search(q.cuda(), k, normalize_bank=nb, normalize_query=nqm, flags=L.FORCE_STREAM)
    kk = min(k, n)
    truth = ss.truth_scores_fp64(q, bank, normalize_bank=bool(nb), normalize_query=bool(nqm))
    ss.compare_topk(a[0].cpu().numpy()[:, :kk], a[1].cpu().numpy()[:, :kk], truth, k, gap=GAP, rtol=RTOL, atol=ATOL)
    assert bool((a[1][:, kk:] == -1).all())
    if nfb == 0:
        assert torch.equal(a[1], b[1]) and torch.equal(a[0], b[0]), (n, d, nq, k, nb, nqm)
    fb.close()


@pytest.mark.parametrize("n,d,nq,k", [(3000, 37, 9, 5), (20000, 101, 70, 16), (2500, 385, 33, 32), (129, 3, 6, 2)])
def test_filter_dims_not_multiple_of_four(n, d, nq, k):
    """Rows of a FactBank are padded to 16 bytes, so the tensor path still applies; the exact re-scoring then has to
    follow the stream kernel's SCALAR column order (one column per lane and chunk) to stay bit-identical."""
    bank, q = _randn(n, d, 5 * n + d, unit=False), _randn(nq, d, 3 * nq + d, unit=False)
    fb = FactBank.from_embeddings(bank)
    a = fb.search(q.cuda(), k, flags=FILTER)
    b = fb.search(q.cuda(), k, flags=L.FORCE_STREAM)
    truth = ss.truth_scores_fp64(q, bank)
    kk = min(k, n)
    ss.compare_topk(a[0].cpu().numpy()[:, :kk], a[1].cpu().numpy()[:, :kk], truth, k, gap=GAP, rtol=RTOL, atol=ATOL)
    assert torch.equal(a[1], b[1]) and torch.equal(a[0], b[0])
    fb.close()


@pytest.mark.parametrize("n,d,nq,k,nb,nqm", [(5000, 768, 64, 32, "eps", "eps"), (40000, 256, 300, 16, "eps", "eps"),
                                               (3000, 384, 20, 5, None, None), (20000, 100, 70, 64, "noeps", "noeps"),
                                               (60000, 512, 130, 32, "eps", None), (129, 72, 9, 8, "eps", "eps")])
def test_bf16_bank_filter_matches_the_stream_kernel_bit_for_bit(n, d, nq, k, nb, nqm):
    """bf16 banks take the same filter (queries rounded to bf16 on the tensor core, exact re-scoring against the
    bf16 rows in the stream kernel's arithmetic): identical results, exact w.r.t. the bf16-rounded bank."""
    bank, q = _randn(n, d, 11 * n + d, unit=False), _randn(nq, d, 13 * nq + k, unit=False)
    fb = FactBank.from_embeddings(bank, dtype="bf16")
    _fallbacks()
    a = fb.search(q.cuda(), k, normalize_bank=nb, normalize_query=nqm, flags=FILTER)
    nfb = _fallbacks()
    b = fb.search(q.cuda(), k, normalize_bank=nb, normalize_query=nqm, flags=L.FORCE_STREAM)
    rounded = bank.to(torch.bfloat16).float()
    truth = ss.truth_scores_fp64(q, rounded, normalize_bank=bool(nb), normalize_query=bool(nqm))
    kk = min(k, n)
    ss.compare_topk(a[0].cpu().numpy()[:, :kk], a[1].cpu().numpy()[:, :kk], truth, k, gap=GAP, rtol=RTOL, atol=ATOL)
    assert torch.equal(a[1], b[1]) and torch.equal(a[0], b[0])
    assert nfb <= max(1, nq // 8)
    fb.close()


@pytest.mark.parametrize("n,d,nq,k,nb", [(8174, 4096, 256, 16, "eps"), (2100, 768, 300, 16, "eps"), (18900, 384, 64, 32, None),
                                         (18944, 128, 1100, 5, "eps"), (4100, 256, 33, 32, "noeps"), (6345, 384, 16, 16, "eps")])
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_one_shot_mode_for_short_scans(n, d, nq, k, nb, dtype):
    """Banks of at most one tile per SM (and at least top_k tiles): one tile per CTA, grid-wide threshold exchange,
    candidates straight to a global array (scan_tc.cu oneshot_epilogue).  Bit-identical to the stream kernel and to
    the persistent kernel's list epilogue (DEVQA_DEBUG_NO_ONESHOT), C2's shape included."""
    bank = _hidden_like(n, d, n + d) if d == 4096 else _randn(n, d, n + d, unit=False)
    q = _hidden_like(nq, d, nq + 1) if d == 4096 else _randn(nq, d, nq + 1, unit=False)
    fb = FactBank.from_embeddings(bank, dtype=dtype)
    _fallbacks(); _violations()
    a = fb.search(q.cuda(), k, normalize_bank=nb, normalize_query=nb, flags=FILTER)
    nfb, nv = _fallbacks(), _violations()
    b = fb.search(q.cuda(), k, normalize_bank=nb, normalize_query=nb, flags=FILTER | L.DEBUG_NO_ONESHOT)
    _fallbacks()
    if d <= 1024:
        c = fb.search(q.cuda(), k, normalize_bank=nb, normalize_query=nb, flags=L.FORCE_STREAM)
        assert torch.equal(a[1], c[1]) and torch.equal(a[0], c[0])
    assert torch.equal(a[1], b[1]) and torch.equal(a[0], b[0])
    assert nv == 0 and nfb <= max(1, nq // 16)
    stored = bank.to(torch.bfloat16).float() if dtype == "bf16" else bank
    truth = ss.truth_scores_fp64(q, stored, normalize_bank=bool(nb), normalize_query=bool(nb))
    ss.compare_topk(a[0].cpu().numpy(), a[1].cpu().numpy(), truth, k, gap=GAP, rtol=RTOL, atol=ATOL)
    fb.close()


def test_one_shot_mode_overflow_takes_the_fallback():
    """NaN rows in >= k tiles leave no usable threshold: every row is a candidate, the array overflows, the query is
    flagged and the exact pipeline answers it."""
    bank = _randn(5000, 128, 3, unit=False)
    bank[::100, 7] = float("nan")                      # 50 NaN rows spread over all 40 tiles
    q = _randn(12, 128, 4, unit=False)
    fb = FactBank.from_embeddings(bank)
    _fallbacks()
    s, i = fb.search(q.cuda(), 8, flags=FILTER)
    assert _fallbacks() >= 1
    s2, i2 = fb.search(q.cuda(), 8, flags=L.FORCE_STREAM)
    assert torch.isnan(s).all() and torch.equal(i, i2)
    fb.close()


@pytest.mark.parametrize("n,d,nq,k,nb", [(30000, 256, 40, 65, "eps"), (50000, 128, 70, 100, None), (30000, 768, 20, 128, "eps"),
                                         (3000, 384, 33, 128, "noeps"), (100, 64, 9, 128, "eps")])
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_filter_serves_top_k_up_to_128(n, d, nq, k, nb, dtype):
    """top_k in 65..128 used to leave the filter for the 3xTF32 scan.  Lists of capacity 128,
    K' = 256 merged candidates, refine_kernel over 256 slots: bit-identical to the stream kernel."""
    bank, q = _randn(n, d, n + k, unit=False), _randn(nq, d, nq + k, unit=False)
    fb = FactBank.from_embeddings(bank, dtype=dtype)
    _fallbacks(); _violations()
    a = fb.search(q.cuda(), k, normalize_bank=nb, normalize_query=nb, flags=FILTER)
    nfb, nv = _fallbacks(), _violations()
    b = fb.search(q.cuda(), k, normalize_bank=nb, normalize_query=nb, flags=L.FORCE_STREAM)
    assert nv == 0 and nfb <= max(1, nq // 8)
    assert torch.equal(a[1], b[1]) and torch.equal(a[0], b[0])
    kk = min(k, n)
    stored = bank.to(torch.bfloat16).float() if dtype == "bf16" else bank
    truth = ss.truth_scores_fp64(q, stored, normalize_bank=bool(nb), normalize_query=bool(nb))
    ss.compare_topk(a[0].cpu().numpy()[:, :kk], a[1].cpu().numpy()[:, :kk], truth, k, gap=GAP, rtol=RTOL, atol=ATOL)
    assert bool((a[1][:, kk:] == -1).all())
    fb.close()


# ---- bf16 shadow of an fp32 bank: exact fp32 results from half the bytes --------------------------------------------
@pytest.mark.parametrize("n,d,nq,k,nb,unit", [(40000, 768, 64, 32, "eps", True), (25000, 384, 5, 16, "eps", False),
                                              (60000, 128, 130, 5, None, False), (19000, 100, 70, 64, "noeps", False),
                                              (300000, 64, 1, 32, "eps", True), (20000, 768, 300, 32, None, True)])
def test_shadow_bank_returns_the_exact_fp32_results(n, d, nq, k, nb, unit):
    """The filter scans a scaled-fp16 SHADOW of the fp32 bank (2 bytes per element), candidates are re-scored against
    the fp32 rows: bit-identical to the stream kernel on the fp32 bank, canary silent."""
    bank, q = _randn(n, d, n + d, unit=unit), _randn(nq, d, nq + k, unit=unit)
    if not unit:
        bank = bank * torch.exp(0.5 * torch.randn(n, 1, generator=torch.Generator().manual_seed(n)))
    fb = FactBank.from_embeddings(bank, shadow=True)
    assert fb.has_shadow
    _fallbacks(); _violations()
    a = fb.search(q.cuda(), k, normalize_bank=nb, normalize_query=nb, flags=L.FORCE_TENSOR)
    nfb, nv = _fallbacks(), _violations()
    b = fb.search(q.cuda(), k, normalize_bank=nb, normalize_query=nb, flags=L.FORCE_STREAM)
    c = fb.search(q.cuda(), k, normalize_bank=nb, normalize_query=nb, flags=L.FORCE_TENSOR | L.NO_SHADOW)
    assert nv == 0 and nfb <= max(1, nq // 8), (nfb, nv)
    assert torch.equal(a[1], b[1]) and torch.equal(a[0], b[0])
    assert torch.equal(a[1], c[1]) and torch.equal(a[0], c[0])
    truth = ss.truth_scores_fp64(q, bank, normalize_bank=bool(nb), normalize_query=bool(nb))
    ss.compare_topk(a[0].cpu().numpy(), a[1].cpu().numpy(), truth, k, gap=GAP, rtol=RTOL, atol=ATOL)
    # appended rows reach the shadow too; dropping the shadow changes nothing
    extra = _randn(1000, d, 5, unit=unit) * 3.0
    fb.append(extra)
    a2 = fb.search(q.cuda(), k, normalize_bank=nb, normalize_query=nb, flags=L.FORCE_TENSOR)
    b2 = fb.search(q.cuda(), k, normalize_bank=nb, normalize_query=nb, flags=L.FORCE_STREAM)
    assert torch.equal(a2[1], b2[1]) and torch.equal(a2[0], b2[0])
    assert bool((a2[1] >= n).any()) or nb is not None                # (tripled rows show up in un-normalised searches)
    fb.set_shadow(False)
    assert not fb.has_shadow
    a3 = fb.search(q.cuda(), k, normalize_bank=nb, normalize_query=nb, flags=L.FORCE_TENSOR)
    assert torch.equal(a3[1], b2[1]) and torch.equal(a3[0], b2[0])
    fb.close()


@pytest.mark.parametrize("d", [768, 4096])
def test_shadow_error_bound_on_the_constructed_worst_case(d):
    """The shadow is fp16 with a power-of-two scale per row / per query (11 significand bits).  Rows AND queries one
    below the fp16 rounding midpoint, parallel, all positive: the per-element error reaches 2^-10 of ||row|| ||query||;
    the shadow's bound (2^-10 + 2^-22 + 2^-19 + D 2^-22) must hold, and the real search must come back exact with the
    canary silent -- also for rows scaled far outside fp16's own range."""
    rng = np.random.default_rng(d)
    n, nq = 25000, 16                                         # >= 148 tiles: the shadow path is taken
    e = rng.integers(120, 128, size=(n, d), dtype=np.uint32)
    rows = _bits_to_f32((e << np.uint32(23)) | np.uint32(0x0FFF))
    queries = _bits_to_f32((e[:nq] << np.uint32(23)) | np.uint32(0x0FFF))
    big = torch.exp2(torch.randint(-40, 40, (n, 1), generator=torch.Generator().manual_seed(d)).float())   # (||row||^2 stays in fp32 range)
    rows = rows * big                                          # exact (powers of two): fp16 alone could not hold these
    fb = FactBank.from_embeddings(rows, shadow=True)
    rel = 2.0 ** -10 + 2.0 ** -22 + 2.0 ** -19 + d * 2.0 ** -22
    _fallbacks(); _violations()
    a = fb.search(queries.cuda(), 8, normalize_bank="eps", normalize_query=None, flags=L.FORCE_TENSOR)
    nfb, nv = _fallbacks(), _violations()
    b = fb.search(queries.cuda(), 8, normalize_bank="eps", normalize_query=None, flags=L.FORCE_STREAM)
    assert nv == 0
    assert torch.equal(a[1], b[1]) and torch.equal(a[0], b[0])
    assert a[1][:, 0].cpu().tolist() == list(range(nq))        # query j is parallel to row j: cosine 1
    # the measured error of the fp16 product for the parallel pairs (scales are exact powers of two), against the bound
    r0 = rows[:nq] / big[:nq]
    r16, q16 = r0.half().double(), queries.half().double()
    exact = (r0.double() * queries.double()).sum(1)
    approx = (r16 * q16).sum(1)
    scale = r0.double().norm(dim=1) * queries.double().norm(dim=1)
    err = float(((exact - approx).abs() / scale).max())
    _record(f"shadow_fp16_worstcase_d{d}", {"max_err": err, "bound": rel, "fallbacks": nfb})
    assert 0.9 * 2.0 ** -10 < err < rel
    # un-normalised scores over the same bank (E scales with ||row||: rows up to 2^40 dominate) stay exact too
    c = fb.search(queries.cuda(), 8, normalize_bank=None, normalize_query=None, flags=L.FORCE_TENSOR)
    dd = fb.search(queries.cuda(), 8, normalize_bank=None, normalize_query=None, flags=L.FORCE_STREAM)
    assert _violations() == 0
    assert torch.equal(c[1], dd[1]) and torch.equal(c[0], dd[0])
    fb.close()


@pytest.mark.parametrize("n,d,nq,k,nb,shadow", [(40000, 768, 64, 32, "eps", True), (40000, 768, 64, 32, "eps", False),
                                                (150000, 128, 300, 64, None, True), (19000, 384, 9, 5, "noeps", False),
                                                (700000, 64, 1030, 16, "eps", True)])
def test_emit_mode_equals_the_list_epilogue(n, d, nq, k, nb, shadow):
    """Long filter scans append their survivors to a flat candidate array (emit mode) instead of keeping per-CTA
    sorted lists: same results as the list epilogue (DEVQA_DEBUG_NO_EMIT) and as the stream kernel, bit for bit."""
    bank, q = _randn(n, d, n + d, unit=False), _randn(nq, d, nq + d, unit=False)
    fb = FactBank.from_embeddings(bank, shadow=shadow)
    _fallbacks(); _violations()
    a = fb.search(q.cuda(), k, normalize_bank=nb, normalize_query=nb, flags=L.FORCE_TENSOR)
    nfb, nv = _fallbacks(), _violations()
    b = fb.search(q.cuda(), k, normalize_bank=nb, normalize_query=nb, flags=L.FORCE_TENSOR | L.DEBUG_NO_EMIT)
    _fallbacks()
    c = fb.search(q[:8].cuda(), k, normalize_bank=nb, normalize_query=nb, flags=L.FORCE_STREAM)
    assert nv == 0 and nfb <= max(1, nq // 16), (nfb, nv)
    assert torch.equal(a[1], b[1]) and torch.equal(a[0], b[0])
    assert torch.equal(a[1][:8], c[1]) and torch.equal(a[0][:8], c[0])
    fb.close()


@pytest.mark.parametrize("n,d,nq,k,shadow", [(60000, 768, 300, 32, True), (45000, 768, 64, 32, True), (60000, 384, 130, 16, False)])
def test_scan_variants_are_bit_identical(n, d, nq, k, shadow):
    """The scan kernel's alternatives behind debug bits -- one group of epilogue warps instead of two (bit 4), bank tiles
    without the L2 evict-first hint (5), query operand streamed with the stages instead of resident (6), CTA pairs also at
    64 / 128 queries per pass (29) -- change how the tiles are multiplied and scanned, never what comes back."""
    bank, q = _randn(n, d, 3 * n + d, unit=False), _randn(nq, d, 5 * nq + d, unit=False)
    fb = FactBank.from_embeddings(bank, shadow=shadow)
    _fallbacks(); _violations()
    ref = fb.search(q.cuda(), k, normalize_bank="eps", normalize_query="eps", flags=L.FORCE_TENSOR)
    for bit in (4, 5, 6, 29):
        out = fb.search(q.cuda(), k, normalize_bank="eps", normalize_query="eps", flags=L.FORCE_TENSOR | (1 << bit))
        assert torch.equal(out[1], ref[1]) and torch.equal(out[0], ref[0]), bit
    assert _violations() == 0 and _fallbacks() <= 5 * max(1, nq // 16)
    c = fb.search(q[:8].cuda(), k, normalize_bank="eps", normalize_query="eps", flags=L.FORCE_STREAM)
    assert torch.equal(ref[1][:8], c[1]) and torch.equal(ref[0][:8], c[0])
    fb.close()


def test_emit_mode_overflow_takes_the_fallback():
    """20000 identical rows: every one of them beats any threshold, the candidate array (8192 per query) overflows,
    the query is flagged and the exact pipeline answers it -- ties in ascending id order."""
    bank = _randn(60000, 64, 5)
    bank[10000:30000] = bank[10000]
    q = torch.cat([bank[10000:10001], _randn(7, 64, 6)])
    fb = FactBank.from_embeddings(bank)
    _fallbacks()
    s, i = fb.search(q.cuda(), 16, flags=L.FORCE_TENSOR)
    assert _fallbacks() >= 1
    assert i[0].tolist() == list(range(10000, 10016))
    s2, i2 = fb.search(q.cuda(), 16, flags=L.FORCE_STREAM)
    assert torch.equal(i, i2) and torch.equal(s, s2)
    fb.close()


@pytest.mark.parametrize("shadow", [False, True])
def test_emit_mode_with_nan_inf_and_zero_rows(shadow):
    """Long scan (emit mode, with and without the bf16 shadow) over a bank with a zero row, NaN rows, an inf entry and
    a row far beyond bf16's precision needs: NaN ranks first like torch.topk, everything else exact."""
    bank = _randn(40000, 128, 91, unit=False)
    bank[17] = 0
    bank[99, 5] = float("nan")
    bank[30000, 7] = float("nan")
    bank[555] *= 1e30
    bank[777] *= 1e-30
    q = _randn(12, 128, 92, unit=False)
    fb = FactBank.from_embeddings(bank, shadow=shadow)
    for nb in ("eps", "noeps", None):
        a = fb.search(q.cuda(), 8, normalize_bank=nb, normalize_query=nb, flags=L.FORCE_TENSOR)
        b = fb.search(q.cuda(), 8, normalize_bank=nb, normalize_query=nb, flags=L.FORCE_STREAM)
        assert torch.equal(a[1], b[1]), nb
        assert torch.equal(torch.nan_to_num(a[0], nan=7.0), torch.nan_to_num(b[0], nan=7.0)), nb
        nan_first = {99, 30000} | ({17} if nb == "noeps" else set())
        assert set(a[1][0, :len(nan_first)].tolist()) == nan_first and bool(torch.isnan(a[0][:, 0]).all())
    fb.close()
