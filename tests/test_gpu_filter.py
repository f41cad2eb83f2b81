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


def relerr(d):
    """csrc/scan.cuh tc_filter_relerr"""
    return 1.5 * 2.0 ** -10 + d * 2.0 ** -24


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


def test_duplicated_rows_take_the_exact_fallback():
    """200 identical consecutive rows: one CTA's list fills up with ties, rows it turns away are not provably
    irrelevant, and the 3xTF32 scan redoes that query; ties resolve to ascending ids either way."""
    bank = _randn(5000, 256, 11)
    bank[1000:1200] = bank[1000]
    q = torch.cat([bank[1000:1001] + 0.01 * _randn(1, 256, 12), _randn(15, 256, 13)])
    fb = FactBank.from_embeddings(bank)
    _fallbacks()
    s, i = fb.search(q.cuda(), 32, flags=FILTER)
    n_fb = _fallbacks()
    assert 1 <= n_fb <= 2
    assert i[0].tolist() == list(range(1000, 1032))
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
    bank = _randn(30000, 128, 61)
    bank[5000:5400] = bank[5000]
    q = torch.cat([bank[5000:5001] + 0.01 * _randn(1, 128, 62), _randn(31, 128, 63)])
    fb = FactBank.from_embeddings(bank)
    _fallbacks()
    s, i = fb.search(q.cuda(), 32, flags=FILTER)
    assert _fallbacks() >= 1
    assert i[0].tolist() == list(range(5000, 5032))
    ss.compare_topk(s.cpu().numpy(), i.cpu().numpy(), ss.truth_scores_fp64(q, bank), 32, gap=GAP, rtol=RTOL, atol=ATOL)
    s2, i2 = fb.search(q.cuda(), 32, flags=L.FORCE_STREAM)
    assert torch.equal(i, i2) and torch.equal(s, s2)
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
