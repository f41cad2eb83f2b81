"""Parity of the CUDA path (through the C ABI) against the CPU oracle.  GPU only.

Bar (BASELINE.json north_star): identical top-k indices wherever the fp64 score gap exceeds
1e-5; fp32 scores within 1e-5 relative (plus 1e-6 absolute for scores near zero); bf16-bank mode
is reported as recall@k, not held to the fp32 bar.
"""
import numpy as np
import pytest
import torch

import devqa_b200
from devqa_b200 import FactBank, util as dutil
from oracle import semantic_search as ss
from oracle.st_shim import StubEncoder

pytestmark = pytest.mark.gpu

GAP, RTOL, ATOL = 1e-5, 1e-5, 1e-6
L = devqa_b200._lib


def _randn(n, d, seed, unit=True):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, d, generator=g)
    return torch.nn.functional.normalize(x, dim=1) if unit else x


def _hidden_like(n, d, seed):
    """C2-style un-normalised rows: randn * exp(N(0, 0.5^2)) per row + 4 outlier dims x50."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, d, generator=g) * torch.exp(0.5 * torch.randn(n, 1, generator=g))
    x[:, [1, d // 3, d // 2, d - 2]] *= 50.0
    return x


def _check(bank_cpu, q_cpu, k, nb="eps", nq="eps", dtype="fp32", flags=0):
    bank = FactBank.from_embeddings(bank_cpu, dtype=dtype)
    scores, ids = bank.search(q_cpu.cuda(), k, normalize_bank=nb, normalize_query=nq, flags=flags)
    scores, ids = scores.cpu().numpy(), ids.cpu().numpy()
    kk = min(k, len(bank_cpu))
    assert np.all(ids[:, kk:] == -1) and np.all(np.isneginf(scores[:, kk:]))
    truth = ss.truth_scores_fp64(q_cpu, bank_cpu, normalize_bank=bool(nb), normalize_query=bool(nq))
    stats = ss.compare_topk(scores[:, :kk], ids[:, :kk], truth, k, gap=GAP, rtol=RTOL, atol=ATOL)
    # and against the reference composition itself (torch-CPU oracle), same gap rule
    o_s, o_i = ss.search_arrays(q_cpu, bank_cpu, k, nb or "none", nq or "none")
    ss.compare_topk(o_s, o_i, truth, k, gap=GAP, rtol=RTOL, atol=ATOL)
    bank.close()
    return stats


# ---- BASELINE configs at oracle-friendly sizes ------------------------------------------------------------
def test_c1_evqa_shape_single_queries():
    """C1: 6345 x 384 unit-norm MiniLM-like bank, one query at a time, k=16 and tops=5."""
    bank = _randn(6345, 384, 1234)
    g = torch.Generator().manual_seed(5678)
    near = torch.nn.functional.normalize(bank[:16] + 0.1 * torch.randn(16, 384, generator=g), dim=1)
    far = _randn(16, 384, 5679)
    fb = FactBank.from_embeddings(bank)
    for q in torch.cat([near, far]):
        for k in (16, 5):
            s, i = fb.search(q.cuda(), k)
            truth = ss.truth_scores_fp64(q[None], bank)
            ss.compare_topk(s.cpu().numpy(), i.cpu().numpy(), truth, k, gap=GAP, rtol=RTOL, atol=ATOL)
    s, i = fb.search(near.cuda(), 1)
    assert i[:, 0].cpu().tolist() == list(range(16))              # noisy copies retrieve their source row
    fb.close()


def test_c2_vlkeb_shape_unnormalised_rows():
    """C2 (row-reduced): D=4096 hidden-state-like rows, fused normalise ON, Q=64, k=16."""
    stats = _check(_hidden_like(2048, 4096, 2234), _hidden_like(64, 4096, 6678), 16)
    assert stats["pinned_positions"] > 900


def test_c3_shape_all_query_tiles():
    """C3 (row-reduced): D=768, k=32, Q in {1,2,3,4,7,8,9,16,17,64}: every QT variant + pass tails."""
    bank = _randn(20011, 768, 3234)
    for nq in (1, 2, 3, 4, 7, 8, 9, 16, 17, 64):
        _check(bank, _randn(nq, 768, 7678 + nq), 32)


def test_query_chunk_edges():
    """Q = 100 / 101: the reference's query_chunk_size edge (invisible in results)."""
    bank = _randn(3000, 128, 21)
    _check(bank, _randn(100, 128, 22), 10)
    _check(bank, _randn(101, 128, 23), 10)


# ---- edge cases the domain has -------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 5, 31, 32, 33, 127, 129, 1000, 4097])
def test_ragged_bank_sizes(n):
    _check(_randn(n, 96, 100 + n), _randn(3, 96, 200 + n), 16)


@pytest.mark.parametrize("k", [1, 2, 5, 16, 31, 32, 33, 64, 100, 128, 500, 1024])
def test_top_k_range(k):
    _check(_randn(3001, 64, 300 + k), _randn(2, 64, 400 + k), k)


@pytest.mark.parametrize("d", [1, 3, 4, 5, 63, 64, 127, 130, 384, 385, 1000])
def test_odd_dims_take_the_scalar_path(d):
    _check(_randn(777, d, 500 + d, unit=False), _randn(4, d, 600 + d, unit=False), 8)


def test_k_larger_than_bank_and_too_large_k():
    bank = _randn(7, 32, 31)
    fb = FactBank.from_embeddings(bank)
    s, i = fb.search(bank[:2].cuda(), 16)
    assert (i[:, 7:] == -1).all() and sorted(i[0, :7].cpu().tolist()) == list(range(7))
    hits = fb.hits(bank[:2], 16)
    assert [len(h) for h in hits] == [7, 7] and hits[0][0]["corpus_id"] == 0 and hits[1][0]["corpus_id"] == 1
    s, i = fb.search(bank[:1].cuda(), 1025)                      # any top_k, like upstream: min(top_k, N) results
    assert sorted(i[0, :7].cpu().tolist()) == list(range(7)) and bool((i[0, 7:] == -1).all())
    with pytest.raises(ValueError):
        fb.search(bank[:1].cuda(), 0)
    fb.close()


def test_empty_bank_and_empty_query_batch():
    fb = FactBank(16)
    s, i = fb.search(torch.zeros(2, 16).cuda(), 4)
    assert (i == -1).all() and torch.isneginf(s).all()
    fb.append(torch.ones(3, 16))
    s, i = fb.search(torch.zeros(0, 16).cuda(), 4)
    assert s.shape == (0, 4)
    fb.close()


def test_zero_rows_eps_clamp_and_noeps_nan():
    """F.normalize semantics: a zero row scores 0 (eps clamp).  HICE semantics (no eps): its score
    is NaN and ranks FIRST, as torch.topk does in the reference."""
    bank = _randn(200, 64, 41)
    bank[17] = 0
    q = _randn(3, 64, 42)
    fb = FactBank.from_embeddings(bank)
    s, i = fb.search(q.cuda(), 200)
    s, i = s.cpu().numpy(), i.cpu().numpy()
    for r in range(3):
        assert s[r][i[r] == 17] == 0.0
    s, i = fb.search(q.cuda(), 5, normalize_bank="noeps", normalize_query="noeps")
    assert (i[:, 0] == 17).all() and torch.isnan(s[:, 0]).all() and not torch.isnan(s[:, 1:]).any()
    # The reference's own NaN order is implementation-defined (torch.topk lifts NaN, the Python heap
    # then mis-orders it: SURVEY.md Appendix A), so the oracle is consulted on the non-NaN rows only.
    keep = [r for r in range(200) if r != 17]
    o_s, o_i = ss.search_arrays(q, bank[keep], 4, "noeps", "noeps")
    np.testing.assert_array_equal(i[:, 1:].cpu().numpy(), np.asarray(keep)[o_i])
    fb.close()


def test_exact_ties_are_ordered_by_ascending_id():
    """Duplicated rows: the reference's tie order is implementation-defined; ours is id-ascending."""
    base = _randn(50, 48, 51)
    bank = torch.cat([base, base[:10], base[:10]])
    fb = FactBank.from_embeddings(bank)
    s, i = fb.search(base[3].cuda(), 3)
    assert i[0].cpu().tolist() == [3, 53, 63] and s[0, 0] == s[0, 1] == s[0, 2]
    fb.close()


def test_large_norm_outlier_rows():
    bank = _hidden_like(3000, 256, 61)
    bank[5] *= 1e4
    bank[6] *= 1e-4
    _check(bank, _hidden_like(8, 256, 62), 16)


def test_dot_score_without_normalisation():
    _check(_randn(1500, 72, 71, unit=False), _randn(5, 72, 72, unit=False), 12, nb=None, nq=None)


def test_noeps_matches_eps_on_nonzero_rows():
    _check(_randn(1500, 72, 73, unit=False), _randn(5, 72, 74, unit=False), 12, nb="noeps", nq="noeps")


def test_bank_only_normalisation():
    _check(_randn(900, 40, 75, unit=False), _randn(5, 40, 76, unit=False), 7, nb="eps", nq=None)


# ---- the util shim: reference call pattern -----------------------------------------------------------------
def test_semantic_search_shim_matches_oracle_format():
    bank = _randn(5000, 384, 81, unit=False)
    q = _randn(3, 384, 82, unit=False)
    nb = dutil.normalize_embeddings(bank.cuda())
    nq = dutil.normalize_embeddings(q.cuda())
    np.testing.assert_allclose(nb.cpu().numpy(), ss.normalize_embeddings(bank).numpy(), rtol=1e-6, atol=1e-9)
    hits = dutil.semantic_search(nq, nb, score_function=dutil.dot_score, top_k=5)
    want = ss.semantic_search(ss.normalize_embeddings(q), ss.normalize_embeddings(bank), score_function=ss.dot_score, top_k=5)
    assert [[h["corpus_id"] for h in r] for r in hits] == [[h["corpus_id"] for h in r] for r in want]
    for r, w in zip(hits, want):
        np.testing.assert_allclose([h["score"] for h in r], [h["score"] for h in w], rtol=RTOL)
        assert set(r[0]) == {"corpus_id", "score"} and isinstance(r[0]["corpus_id"], int)
    # default score function = cos_sim; 1-D query; CPU inputs are uploaded; list length min(k, N)
    h2 = dutil.semantic_search(q[0], bank, top_k=5)
    assert [h["corpus_id"] for h in h2[0]] == [h["corpus_id"] for h in want[0]]
    assert len(dutil.semantic_search(q, bank[:3], top_k=10)[0]) == 3


def test_shim_on_tensor_views_and_slices():
    """Caller-owned banks arrive as slices / strided views (e.g. corpus[c0:c1])."""
    big = _randn(4000, 100, 91)
    cuda = big.cuda()
    view = cuda[1001:3001]                   # row offset: still 16-byte aligned rows (D % 4 == 0)
    s, i = dutil.semantic_search_arrays(big[:4].cuda(), view, 6, score_function=dutil.dot_score)
    truth = ss.truth_scores_fp64(big[:4], big[1001:3001], False, False)
    ss.compare_topk(s.cpu().numpy(), i.cpu().numpy(), truth, 6, gap=GAP, rtol=RTOL, atol=ATOL)
    cols = cuda[:, 3:99]                     # column slice: ld > dim, misaligned base -> scalar path
    s, i = dutil.semantic_search_arrays(big[:4, 3:99].cuda(), cols, 6, score_function=dutil.dot_score)
    truth = ss.truth_scores_fp64(big[:4, 3:99], big[:, 3:99], False, False)
    ss.compare_topk(s.cpu().numpy(), i.cpu().numpy(), truth, 6, gap=GAP, rtol=RTOL, atol=ATOL)


# ---- golden fixtures (reference glue executed in the build container) -------------------------------------------
def test_golden_finds_sim_and_lookups_on_gpu(golden, tmp_path):
    import pickle
    import types
    from devqa_b200 import retrieval
    retrieval.clear_bank_cache()
    stored = {"sentences": golden["bank"]["sentences"], "embeddings": golden["embeddings"],
              "images": golden["bank"]["images"], "prompts": golden["bank"]["prompts"]}
    (tmp_path / "embedding").mkdir()
    for name, keys in [("vqa_embeddings_llava.pkl", stored), ("vqa_embeddings.pkl", {k: stored[k] for k in ("sentences", "embeddings")})]:
        with open(tmp_path / "embedding" / name, "wb") as f:
            pickle.dump(keys, f, protocol=pickle.HIGHEST_PROTOCOL)
    enc = StubEncoder()
    rd = retrieval.RetrievalData()
    rd.init_retrieval(types="vqa", bank_path=str(tmp_path / "embedding" / "vqa_embeddings_llava.pkl"), sentence_model=enc)
    for case in golden["finds_sim"]:
        ex, img = rd.finds_sim(case["src"], case["trg"])
        assert list(ex) == case["example"] and img == case["image"]
        hit = rd._top_hits(case["src"], 5)
        assert [h["corpus_id"] for h in hit] == case["hit_ids"]
        np.testing.assert_allclose([h["score"] for h in hit], case["hit_scores"], rtol=RTOL)
    hp = types.SimpleNamespace(results_dir=str(tmp_path), task_name="vqa", sentence_model_name="x/stub",
                               k=golden["k"], device=0, alg_name="IKE")
    sents = golden["bank"]["sentences"]
    hstored = {"sentences": sents, "embeddings": torch.tensor(golden["embeddings"]).cuda()}
    for case in golden["lookups"]:
        req = {"prompt": case["prompt"], "target": case["target"]}
        ike = retrieval.apply_ike_to_multimodal_model(None, None, req, hp, train_ds=[0], sentence_model=enc)
        assert ike[:-1] == [sents[i] for i in case["ike_ids"]] and ike[-1] == case["ike_tail"]
        hice = retrieval.apply_hice_to_multimodal_model(None, None, req, hp, train_ds=[0], sentence_model=enc,
                                                        stored_data=hstored)
        assert hice[:-1] == [sents[i] for i in case["hice_ids"]]
        sel40 = retrieval.select_icl_samples(case["prompt"], case["target"], hstored, enc, "cuda:0",
                                             types.SimpleNamespace(k=40, device=0))
        assert sel40 == [sents[i] for i in case["select_k40_ids"]]
    retrieval.clear_bank_cache()


# ---- bank object ----------------------------------------------------------------------------------------------
def test_append_is_in_place_and_equivalent_to_bulk_load():
    rows = _randn(5000, 128, 95)
    q = _randn(4, 128, 96)
    bulk = FactBank.from_embeddings(rows)
    inc = FactBank(128)
    pos = 0
    for step in [1, 1, 2, 7, 100, 1, 889, 4000]:
        chunk = rows[pos:pos + step]
        inc.append(chunk.cuda() if pos % 2 else chunk.numpy())          # device and host sources
        pos += step
    assert len(inc) == len(bulk) == 5000 and inc.capacity >= 5000
    torch.testing.assert_close(inc.rows(), rows.cuda(), rtol=0, atol=0)
    a = bulk.search(q.cuda(), 16)
    b = inc.search(q.cuda(), 16)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    inc.truncate(1000)
    s, i = inc.search(q.cuda(), 16)
    assert len(inc) == 1000 and int(i.max()) < 1000
    bulk.close(); inc.close()


def test_host_entry_point_equals_device_entry_point():
    rows, q = _randn(3000, 384, 97), _randn(9, 384, 98)
    fb = FactBank.from_embeddings(rows)
    ds, di = fb.search(q.cuda(), 16)
    hs, hi = fb.search_host(q.numpy(), 16)
    np.testing.assert_array_equal(ds.cpu().numpy(), hs)
    np.testing.assert_array_equal(di.cpu().numpy(), hi)
    fb.close()


def test_bf16_bank_recall_reported_separately():
    rows, q = _randn(20000, 768, 101), _randn(16, 768, 102)
    fb32, fb16 = FactBank.from_embeddings(rows), FactBank.from_embeddings(rows, dtype="bf16")
    _, i32 = fb32.search(q.cuda(), 32)
    s16, i16 = fb16.search(q.cuda(), 32)
    recall = np.mean([len(set(a) & set(b)) / 32 for a, b in zip(i32.cpu().tolist(), i16.cpu().tolist())])
    assert recall >= 0.9, recall
    # the bf16 path is exact with respect to the bf16-rounded bank
    rounded = rows.to(torch.bfloat16).float()
    truth = ss.truth_scores_fp64(q, rounded)
    ss.compare_topk(s16.cpu().numpy(), i16.cpu().numpy(), truth, 32, gap=GAP, rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(fb16.rows(0, 100), rounded[:100].cuda(), rtol=0, atol=0)
    fb32.close(); fb16.close()


def test_launch_counter_moves():
    before = L.lib().devqa_launch_count()
    fb = FactBank.from_embeddings(_randn(100, 32, 111))
    fb.search(_randn(1, 32, 112).cuda(), 4)
    torch.cuda.synchronize()
    assert L.lib().devqa_launch_count() >= before + 2           # scan + merge
    fb.close()


def test_search_is_capturable_in_a_cuda_graph():
    """The tensor API enqueues kernels only (no host synchronisation, the exact-fallback decision is taken on the
    device), so a search can be captured once and replayed: stream kernel (Q = 1) and tensor-core filter path (Q = 40)."""
    import devqa_b200
    from devqa_b200 import FactBank
    g = torch.Generator().manual_seed(77)
    bank = torch.nn.functional.normalize(torch.randn(60_000, 256, generator=g), dim=1)
    fb = FactBank.from_embeddings(bank)
    for nq in (1, 40):
        q = torch.nn.functional.normalize(torch.randn(nq, 256, generator=g), dim=1).cuda()
        q2 = torch.nn.functional.normalize(torch.randn(nq, 256, generator=g), dim=1).cuda()
        want1, want2 = fb.search(q, 16), fb.search(q2, 16)          # also sizes every scratch buffer
        static_q = q.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fb.search(static_q, 16)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = fb.search(static_q, 16)
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(out[0], want1[0]) and torch.equal(out[1], want1[1])
        static_q.copy_(q2)
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(out[0], want2[0]) and torch.equal(out[1], want2[1])
    fb.close()


def test_bank_producer_can_leave_the_bank_resident(tmp_path):
    """N4: encode_ike_facts_multimodal(..., device=0) writes the frozen pickle AND registers the resident bank for it,
    so the first lookup after the build does not re-read and re-upload the file (ike_main.py:187-194)."""
    import pickle
    import types
    from devqa_b200 import retrieval
    retrieval.clear_bank_cache()
    enc = StubEncoder()
    ds = [{"prompt": f"what is item {i}?", "target": f"thing {i}", "rephrase_prompt": f"item {i} is what?",
           "locality_prompt": f"where is place {i}?", "locality_ground_truth": f"city {i}", "image_path": f"a{i}.jpg",
           "rephrase_image_path": f"b{i}.jpg", "locality_image_path": f"c{i}.jpg"} for i in range(50)]
    hp = types.SimpleNamespace(results_dir=str(tmp_path), task_name="vqa", sentence_model_name="x/stub", k=4, device=0,
                               alg_name="IKE")
    stored = retrieval.encode_ike_facts_multimodal(enc, ds, hp, device=0)
    path = tmp_path / "embedding" / "vqa_embeddings_llava.pkl"
    with open(path, "rb") as f:
        on_disk = pickle.load(f)
    assert list(on_disk) == ["sentences", "embeddings", "images", "prompts"] and len(on_disk["sentences"]) == 150
    assert on_disk["embeddings"].dtype == np.float32 and np.array_equal(on_disk["embeddings"], stored["embeddings"])
    bank = retrieval.bank_from_file(str(path), device=0)          # the registered, already resident bank
    assert len(bank) == 150 and bank.sentences == stored["sentences"]
    assert retrieval.bank_from_file(str(path), device=0) is bank
    q = torch.tensor(enc.encode(stored["sentences"][7]))
    s, i = bank.search(q.cuda(), 1)
    assert int(i[0, 0]) == 7
    retrieval.clear_bank_cache()


# ---- Level-0 drop-in: the reference's own util.* calls, replayed ----------------------------------------------------
def test_level0_call_trace_replay():
    """tests/golden/evqa_trace.*: every call the reference's function bodies (init_retrieval, finds_sim,
    apply_ike_to_multimodal_model, apply_hice_to_multimodal_model, select_icl_samples -- executed unmodified by
    make_golden.py) made into ``sentence_transformers.util``, with the exact tensors they passed.  Replayed here
    against ``devqa_b200.util`` on the GPU: same ids bit for bit, scores within 1e-5 -- the zero-source-change
    integration (INTEGRATION.md level 0) exercised through the reference's own arguments."""
    import json, os
    from conftest import GOLDEN
    trace = json.load(open(os.path.join(GOLDEN, "evqa_trace.json")))["calls"]
    arrays = np.load(os.path.join(GOLDEN, "evqa_trace.npz"))
    dev = {}

    def on_gpu(key):
        if key not in dev:
            dev[key] = torch.from_numpy(arrays[key]).cuda()
        return dev[key]

    n_search = n_norm = 0
    sites = set()
    for call in trace:
        sites.add(call["site"])
        if call["fn"] == "normalize_embeddings":
            out = dutil.normalize_embeddings(on_gpu(call["input"]))
            assert out.is_cuda
            torch.testing.assert_close(out.cpu(), torch.from_numpy(arrays[call["output"]]), rtol=1e-6, atol=1e-7)
            n_norm += 1
        else:
            q = on_gpu(call["query"])
            if call["query_ndim"] == 1:
                q = q.reshape(-1)
            fn = {"dot_score": dutil.dot_score, "cos_sim": dutil.cos_sim}[call["score_function"]]
            hits = dutil.semantic_search(q, on_gpu(call["corpus"]), score_function=fn, top_k=call["top_k"])
            want = call["result"]
            assert len(hits) == len(want)
            for got_row, want_row in zip(hits, want):
                assert [h["corpus_id"] for h in got_row] == [w[0] for w in want_row], call["site"]
                np.testing.assert_allclose([h["score"] for h in got_row], [w[1] for w in want_row], rtol=1e-5, atol=1e-6)
            n_search += 1
    assert n_search >= 100 and n_norm >= 80 and len(sites) == 5


def test_level0_dot_score_over_a_prenormalised_tensor_takes_the_filter():
    """The reference's real call pattern (dataset/vllm.py:66-68,104): normalise first, then
    semantic_search(q, bank, score_function=dot_score) over a caller-owned tensor.  The shim computes ||row||^2 once
    per (storage, view, version) so that the tensor path can run its TF32 filter: results bit-identical to the stream
    kernel, in-place edits and views never served from stale norms."""
    import ctypes as C
    n, d, nq, k = 40_000, 512, 64, 32                       # 82 MB: above the cache's minimum size
    bank = dutil.normalize_embeddings(_randn(n, d, 1, unit=False).cuda())
    q = dutil.normalize_embeddings(_randn(nq, d, 2, unit=False).cuda())
    fb_count = C.c_int64()
    L.check(L.lib().devqa_filter_fallbacks(0, C.byref(fb_count)))
    launches0 = L.lib().devqa_launch_count()
    s, i = dutil.semantic_search_arrays(q, bank, top_k=k, score_function=dutil.dot_score)
    first = L.lib().devqa_launch_count() - launches0
    s_again, i_again = dutil.semantic_search_arrays(q, bank, top_k=k, score_function=dutil.dot_score)
    second = L.lib().devqa_launch_count() - launches0 - first
    assert second == first - 2                               # the ||row||^2 pass and the shadow conversion ran once
    (entry,) = dutil._norm_cache.values()
    (hit,) = entry[1].values()
    assert hit[2] is not None and hit[2].dtype == torch.float16 and hit[2].shape == (n, d)   # 313 tiles: shadowed
    s_ref, i_ref = dutil.semantic_search_arrays(q, bank, top_k=k, score_function=dutil.dot_score, flags=L.FORCE_STREAM)
    assert torch.equal(i, i_ref) and torch.equal(s, s_ref) and torch.equal(i_again, i_ref)
    L.check(L.lib().devqa_filter_fallbacks(0, C.byref(fb_count)))
    assert fb_count.value == 0
    # in-place edit: the version counter moves, the norms are recomputed
    bank[:100].mul_(3.0)
    s2, i2 = dutil.semantic_search_arrays(q, bank, top_k=k, score_function=dutil.dot_score)
    s2_ref, i2_ref = dutil.semantic_search_arrays(q, bank, top_k=k, score_function=dutil.dot_score, flags=L.FORCE_STREAM)
    assert torch.equal(i2, i2_ref) and torch.equal(s2, s2_ref)
    assert bool((i2[:, 0] < 100).float().mean() > 0.9)       # the tripled rows now win
    # a view of the same storage is its own entry
    half = bank[n // 2:]
    s3, i3 = dutil.semantic_search_arrays(q, half, top_k=k, score_function=dutil.dot_score)
    s3_ref, i3_ref = dutil.semantic_search_arrays(q, half, top_k=k, score_function=dutil.dot_score, flags=L.FORCE_STREAM)
    assert torch.equal(i3, i3_ref) and torch.equal(s3, s3_ref)
    # hits format through the list API, cos_sim as well
    hits = dutil.semantic_search(q[:3], bank, top_k=5, score_function=dutil.cos_sim)
    assert len(hits) == 3 and len(hits[0]) == 5
    del bank, half
    import gc; gc.collect()
    assert len(dutil._norm_cache) == 0                        # the storage died, its entry went with it


def test_level0_shadow_is_optional_and_changes_nothing(monkeypatch):
    """DEVQA_LEVEL0_SHADOW=0: the shim keeps only the row norms for a caller-owned corpus (TF32 filter over the fp32
    rows).  With the shadow (default, corpora of >= 148 tiles) the same call streams the 2-byte copy: same bits."""
    n, d, nq, k = 60_000, 384, 70, 32
    bank = dutil.normalize_embeddings(_randn(n, d, 21, unit=False).cuda())
    q = dutil.normalize_embeddings(_randn(nq, d, 22, unit=False).cuda())
    monkeypatch.setenv("DEVQA_LEVEL0_SHADOW", "0")
    s0, i0 = dutil.semantic_search_arrays(q, bank, top_k=k, score_function=dutil.dot_score)
    (entry,) = dutil._norm_cache.values()
    (hit,) = entry[1].values()
    assert hit[2] is None
    monkeypatch.delenv("DEVQA_LEVEL0_SHADOW")
    s1, i1 = dutil.semantic_search_arrays(q, bank, top_k=k, score_function=dutil.dot_score)
    assert hit[2] is not None                                # made on the first call that may use it
    s_ref, i_ref = dutil.semantic_search_arrays(q, bank, top_k=k, score_function=dutil.dot_score, flags=L.FORCE_STREAM)
    assert torch.equal(i0, i_ref) and torch.equal(s0, s_ref) and torch.equal(i1, i_ref) and torch.equal(s1, s_ref)
    # one query over a large corpus: served by the tensor path over the shadow as well, same bits as the stream kernel
    big = dutil.normalize_embeddings(_randn(160_000, 384, 23, unit=False).cuda())
    s2, i2 = dutil.semantic_search_arrays(q[:1], big, top_k=k, score_function=dutil.cos_sim)
    s2r, i2r = dutil.semantic_search_arrays(q[:1], big, top_k=k, score_function=dutil.cos_sim, flags=L.FORCE_STREAM)
    assert torch.equal(i2, i2r) and torch.equal(s2, s2r)
    del bank, big, hit, entry
    import gc; gc.collect()
    assert len(dutil._norm_cache) == 0


def test_reference_format_loaders_shadow_large_banks(tmp_path):
    """FactBank.from_stored_data / from_pickle (what retrieval.py's bank_from_file uses for the reference's .pkl files)
    give a large fp32 bank its shadow on their own (auto_shadow); small banks and DEVQA_SHADOW=0 stay as they are.
    Results are the stream kernel's either way."""
    import os, pickle
    emb = _randn(30_000, 256, 31, unit=False)
    stored = {"sentences": [f"s{i}" for i in range(len(emb))], "embeddings": emb.numpy()}
    path = tmp_path / "evqa_embeddings.pkl"
    with open(path, "wb") as f:
        pickle.dump(stored, f)
    fb = FactBank.from_pickle(str(path))
    assert fb.has_shadow
    small = FactBank.from_stored_data({"sentences": [], "embeddings": emb[:5000].numpy()})
    assert not small.has_shadow
    os.environ["DEVQA_SHADOW"] = "0"
    try:
        off = FactBank.from_stored_data(stored)
        assert not off.has_shadow
    finally:
        del os.environ["DEVQA_SHADOW"]
    q = _randn(20, 256, 32, unit=False).cuda()
    a = fb.search(q, 16, normalize_bank="eps", normalize_query="eps")
    b = off.search(q, 16, normalize_bank="eps", normalize_query="eps")
    c = fb.search(q, 16, normalize_bank="eps", normalize_query="eps", flags=L.FORCE_STREAM)
    assert torch.equal(a[1], c[1]) and torch.equal(a[0], c[0]) and torch.equal(b[1], c[1]) and torch.equal(b[0], c[0])
    for x in (fb, small, off):
        x.close()


@pytest.mark.parametrize("n,d,nq,k,dtype", [(5000, 64, 3, 1025, "fp32"), (9000, 128, 2, 3000, "fp32"), (2500, 36, 1, 2500, "bf16"),
                                            (1500, 64, 2, 4000, "fp32")])
def test_any_top_k_like_upstream(n, d, nq, k, dtype):
    """VERDICT r1 missing #6: upstream's semantic_search accepts any top_k (min(top_k, N) results).  Beyond 1024 the
    ranking comes 1024 at a time, each pass the top-1024 below the previous pass's last key: identical to one full
    sort of the stream kernel's scores."""
    bank, q = _randn(n, d, n + d, unit=False), _randn(nq, d, nq + d, unit=False)
    fb = FactBank.from_embeddings(bank, dtype=dtype)
    s, i = fb.search(q.cuda(), k)
    kk = min(k, n)
    assert bool((i[:, kk:] == -1).all()) and bool(torch.isneginf(s[:, kk:]).all())
    # reference ranking: every row's score from the same kernel (1024 at a time is how it is made; compare with the
    # scores of a k = 1024 search and with fp64 truth)
    s1, i1 = fb.search(q.cuda(), 1024)
    assert torch.equal(i[:, :1024], i1) and torch.equal(s[:, :1024], s1)
    assert bool((s[:, :kk - 1] >= s[:, 1:kk]).all())                       # descending throughout
    for r in range(nq):
        assert len(set(i[r, :kk].tolist())) == kk                           # no row twice, none missing when k >= n
    stored = bank.to(torch.bfloat16).float() if dtype == "bf16" else bank
    truth = ss.truth_scores_fp64(q, stored)
    ss.compare_topk(s.cpu().numpy()[:, :kk], i.cpu().numpy()[:, :kk], truth, kk, gap=GAP, rtol=RTOL, atol=ATOL)
    hits = dutil.semantic_search(q[:1].cuda(), stored.cuda(), top_k=k, score_function=dutil.cos_sim) if dtype == "fp32" else None
    if hits is not None:
        assert len(hits[0]) == kk and [h["corpus_id"] for h in hits[0]] == i[0, :kk].tolist()
    fb.close()
