"""Appended-pool retrieval of the lifelong editors (N1) against the reference's torch composition."""
import pytest
import torch
import torch.nn.functional as F

import devqa_b200
from devqa_b200 import EditPool

pytestmark = pytest.mark.gpu


def _g(seed):
    return torch.Generator().manual_seed(seed)


def test_lte_top1_with_threshold_and_incremental_append():
    """lte_vl.py:94-110,145-150: pool grown one row per edit; cosine max + threshold 0.3."""
    d = 384
    pool_ref = torch.zeros(0, d)
    pool = EditPool(d, capacity=4)
    g = _g(1)
    assert pool.lte_retrieve(torch.randn(d, generator=g), 0.3)[0] == [None]
    for step in range(200):
        new = torch.randn(1, d, generator=g)
        pool_ref = torch.cat([pool_ref, new], 0)               # the reference's O(N) append
        pool.append(new)                                       # ours: in place
        if step % 20 == 0 or step == 199:
            near = pool_ref[step // 2] + 0.2 * torch.randn(d, generator=g)
            far = torch.randn(d, generator=g)
            q = torch.stack([near, far])
            t_sim = F.normalize(q, p=2, dim=1) @ F.normalize(pool_ref, p=2, dim=1).T
            v, i = torch.max(t_sim, dim=1)
            want = [int(ix) if float(val) > 0.3 else None for val, ix in zip(v, i)]
            got, sims = pool.lte_retrieve(q, 0.3)
            assert got == want and got[0] == step // 2
            torch.testing.assert_close(sims, v, rtol=1e-5, atol=1e-6)
    pool.restore(10)
    assert len(pool) == 10 and pool.lte_retrieve(pool_ref[150], 0.3)[0][0] in (None, *range(10))


def test_serac_scaled_dot_top1():
    """serac_vl.py:70-71 + modules.py:52-58: sim = rep1 @ pool.T / sqrt(d); torch.max."""
    d = 768 * 3
    g = _g(2)
    reps = torch.randn(500, d, generator=g)
    pool = EditPool(d)
    pool.append(reps)
    q = reps[[7, 311]] * 0.7 + 0.1 * torch.randn(2, d, generator=g)
    sim = torch.einsum('bd,td->bt', q, reps) / (d ** 0.5)
    v, i = torch.max(sim, dim=1)
    gv, gi = pool.serac_retrieve(q)
    assert gi.cpu().tolist() == i.tolist() == [7, 311]
    torch.testing.assert_close(gv.cpu(), v, rtol=1e-5, atol=1e-5)
    assert bool((gv >= 10).all())


def test_recipe_topk_above_prototype():
    """recipe_vl.py:86-92: top retr_top_k rows whose similarity beats the prototype in row 0."""
    d, k = 1024, 5
    g = _g(3)
    knowl = torch.randn(300, d, generator=g)
    knowl[0] = 0.05 * torch.randn(d, generator=g)              # the prototype: near-zero similarity with everything
    pool = EditPool(d)
    pool.append(knowl)
    q = torch.stack([knowl[5] + knowl[9], torch.randn(d, generator=g), -knowl[1:].mean(0) * 50])
    sim = (q @ knowl.T) / d ** 0.5
    proto = sim[:, :1]
    sorted_sim, order = torch.sort(sim, 1, True)
    mask = sorted_sim[:, :k] > proto
    want = torch.split(torch.masked_select(order[:, :k], mask), mask.sum(1).tolist())
    got, (gs, go) = pool.recipe_retrieve(q, k)
    assert [w.tolist() for w in want] == [x.cpu().tolist() for x in got]
    assert set(got[0].cpu().tolist()[:2]) == {5, 9}
    torch.testing.assert_close(gs.cpu(), sorted_sim[:, :k], rtol=1e-5, atol=1e-5)


def test_find_closest_tokens_matches_reference_composition():
    """N2, editor/vllms_for_edit/base.py:187-196 -- including its normalisation of the queries over dim=1 of
    [b, l, d] (the sequence axis)."""
    from devqa_b200 import VocabBank
    g = _g(4)
    v, d, b, l, k = 3000, 512, 2, 7, 5
    matrix = torch.randn(v, d, generator=g)
    emb = matrix[torch.randint(0, v, (b, l), generator=g)] + 0.3 * torch.randn(b, l, d, generator=g)
    e_n = emb / emb.norm(dim=1, keepdim=True)
    m_n = matrix / matrix.norm(dim=1, keepdim=True)
    ref = torch.topk(torch.matmul(e_n, m_n.T), k, dim=-1)
    vb = VocabBank(matrix)
    toks, sim = vb.find_closest_tokens(emb, top_k=k)
    assert toks.shape == (b, l, k) and toks.cpu().tolist() == ref.indices.tolist()
    torch.testing.assert_close(sim.cpu(), ref.values, rtol=1e-5, atol=1e-6)
    vb.close()


@pytest.mark.parametrize("m", [0, 1, 37, 1500])
def test_liveedit_hard_routing_mask(m):
    """N3, liveedit.py:132-143: vis_sim = einsum('bed,med->bme').mean(2) * scale; map = vis_sim > prototype sim."""
    from devqa_b200 import LiveEditPool
    g = _g(5 + m)
    eqe_n, md, scale = 4, 256, 10.0
    pool_ref = torch.zeros(0, eqe_n, md)
    pool = LiveEditPool(eqe_n, md, capacity=2)
    for start in range(0, m, 100):                               # grown edit by edit (here: in small batches)
        new = torch.randn(min(100, m - start), eqe_n, md, generator=g) / md ** 0.5
        pool_ref = torch.cat([pool_ref, new], 0)
        pool.append(new)
    ivr = torch.randn(1, eqe_n, md, generator=g)
    prot = torch.randn(1, eqe_n, md, generator=g) * (0.0 if m == 1500 else 0.05)   # m = 1500: about half pass (> MAX_TOPK)
    vis_sim = torch.einsum('bed,med->bme', ivr, pool_ref).mean(2) * scale
    vis_sim_prot = torch.einsum('bed,bed->be', ivr, prot).mean(1, True) * scale
    safe = (vis_sim - vis_sim_prot).abs()[0] > 1e-5                 # entries whose verdict does not hinge on the last bits
    want = (vis_sim > vis_sim_prot)[0]
    ids, sims = pool.retrieve(ivr, vis_sim_prot, scale)
    got = torch.zeros(m, dtype=torch.bool)
    got[ids.cpu()] = True
    assert torch.equal(got[safe], want[safe])
    assert ids.cpu().tolist() == sorted(ids.cpu().tolist())
    torch.testing.assert_close(sims.cpu(), vis_sim[0][ids.cpu()], rtol=1e-5, atol=1e-5)
    if m == 1500:
        assert len(ids) > 600
        pool.restore(10)
        assert len(pool) == 10 and bool((pool.retrieve(ivr, vis_sim_prot, scale)[0] < 10).all())


def test_liveedit_retrieval_map_without_host_sync():
    """The mask itself (retrieval_map) from a DEVICE threshold; whole pool passing (no k cap)."""
    from devqa_b200 import LiveEditPool
    g = _g(77)
    eqe_n, md, scale, m = 4, 256, 10.0, 5000
    pool_ref = torch.randn(m, eqe_n, md, generator=g) / md ** 0.5
    pool = LiveEditPool(eqe_n, md)
    pool.append(pool_ref)
    ivr = torch.randn(1, eqe_n, md, generator=g)
    vis_sim = torch.einsum('bed,med->bme', ivr, pool_ref).mean(2) * scale
    for thr in (0.0, float(vis_sim.min()) - 1.0, float(vis_sim.max()) + 1.0):
        prot = torch.tensor([[thr]], device="cuda")
        got = pool.retrieval_map(ivr, prot, scale).cpu()
        want = (vis_sim > thr)[0]
        safe = (vis_sim[0] - thr).abs() > 1e-5
        assert got.dtype == torch.bool and got.shape == (m,)
        assert torch.equal(got[safe], want[safe])
        ids, sims = pool.retrieve(ivr, prot, scale)
        assert torch.equal(torch.nonzero(got)[:, 0], ids.cpu())           # ascending ids == pool[retrieval_map] order
    assert len(pool.retrieve(ivr, torch.tensor([[float(vis_sim.min()) - 1.0]]), scale)[0]) == m


@pytest.mark.parametrize("n,d,nq,dtype,nb", [(1, 64, 1, "fp32", None), (7, 36, 3, "fp32", "eps"), (3001, 384, 5, "fp32", "eps"),
                                            (20000, 768, 9, "bf16", "noeps"), (70000, 101, 2, "fp32", None)])
def test_select_above_matches_the_search_scores_bit_for_bit(n, d, nq, dtype, nb):
    """devqa_bank_select_above: per-query device thresholds, cap smaller than the count, scores identical to
    devqa_bank_search's (the stream kernel's arithmetic), NaN rows never selected."""
    import ctypes as C
    from devqa_b200 import FactBank, norm_flags
    L = devqa_b200._lib
    g = _g(n + d)
    bank = torch.randn(n, d, generator=g)
    if n > 100:
        bank[5, 0] = float("nan")
    q = torch.randn(nq, d, generator=g)
    fb = FactBank.from_embeddings(bank, dtype=dtype)
    k = min(n, 64)
    s_ref, i_ref = fb.search(q.cuda(), k, normalize_bank=nb, normalize_query=nb, flags=L.FORCE_STREAM)
    finite = torch.where(torch.isnan(s_ref), torch.full_like(s_ref, float("inf")), s_ref)
    # threshold = midway between the 10th and 11th best finite score -> exactly the finite part of the top-10 passes
    fin_sorted = torch.sort(torch.where(torch.isnan(s_ref), torch.full_like(s_ref, -float("inf")), s_ref), dim=1, descending=True)[0]
    cut = min(10, k - 1)
    thr = ((fin_sorted[:, cut - 1] + fin_sorted[:, cut]) / 2 if k > 1 else fin_sorted[:, 0] - 1.0).contiguous()
    cap = 6
    mask = torch.empty((nq, n), dtype=torch.uint8, device="cuda")
    ids = torch.empty((nq, cap), dtype=torch.int64, device="cuda")
    sims = torch.empty((nq, cap), dtype=torch.float32, device="cuda")
    cnt = torch.empty((nq,), dtype=torch.int64, device="cuda")
    qd = q.cuda().contiguous()
    L.check(L.lib().devqa_bank_select_above(fb._h, qd.data_ptr(), nq, d, thr.data_ptr(), 1.0, norm_flags(nb, nb),
                                            mask.data_ptr(), ids.data_ptr(), sims.data_ptr(), cap, cnt.data_ptr(),
                                            torch.cuda.current_stream().cuda_stream))
    for j in range(nq):
        want = sorted(int(i) for s, i in zip(s_ref[j].tolist(), i_ref[j].tolist()) if s == s and s > float(thr[j]))
        assert int(cnt[j]) == len(want) == int(mask[j].sum())
        assert torch.nonzero(mask[j])[:, 0].tolist() == want
        assert ids[j].tolist() == (want + [-1] * cap)[:cap]
        lookup = dict(zip(i_ref[j].tolist(), s_ref[j].tolist()))
        for c in range(min(cap, len(want))):
            assert float(sims[j, c]) == lookup[want[c]]                  # bit-identical to the search's score
    fb.close()


def test_hice_memory_gate():
    """HICE.py:238-250: sim = q/||q|| @ (mem/||mem||).T; in scope iff sim.max() > 0.81."""
    g = _g(9)
    d = 384
    mem = [torch.randn(1, d, generator=g) for _ in range(40)]
    pool = EditPool(d)
    assert pool.hice_memory_gate(torch.randn(1, d, generator=g)) == (False, float("-inf"))
    for m in mem:
        pool.append(m)
    memory_fea = torch.cat(mem, dim=0)
    memory_fea = memory_fea / torch.norm(memory_fea, p=2, dim=-1).unsqueeze(-1)
    for q in (mem[17] + 0.3 * torch.randn(1, d, generator=g), torch.randn(1, d, generator=g)):
        q_fea = q / torch.norm(q, p=2, dim=-1)
        sim = q_fea @ memory_fea.T
        gate, best = pool.hice_memory_gate(q)
        assert gate == bool(sim.max() > 0.81)
        assert abs(best - float(sim.max())) < 1e-5
    assert pool.hice_memory_gate(mem[17])[0] is True


# ---- N4: HICE scope gate + memory check + lookup in one call -------------------------------------------------------
def _ranpac_reference(q, W_rand, Wo):
    """multimodal_evaluate.py:387-391"""
    query_fea = q / torch.norm(q, p=2, dim=-1, keepdim=True)
    query_fea_h = torch.nn.functional.relu(query_fea @ W_rand)
    logits = query_fea_h @ Wo.T
    return logits, torch.max(logits, dim=-1)[1]


@pytest.mark.parametrize("d,M,C,n", [(384, 10000, 2, 11), (384, 1000, 2, 1), (100, 777, 3, 40), (768, 4096, 2, 17)])
def test_scope_gate_matches_the_torch_composition(d, M, C, n):
    from devqa_b200 import ScopeGate
    g = _g(d + M)
    W_rand = torch.randn(d, M, generator=g)
    Wo = torch.randn(C, M, generator=g) / M ** 0.5
    q = torch.randn(n, d, generator=g) * 3.0
    gate = ScopeGate(Wo, W_rand)
    logits, preds = gate.predict(q)
    ref_logits, ref_preds = _ranpac_reference(q.double(), W_rand.double(), Wo.double())
    torch.testing.assert_close(logits.cpu().double(), ref_logits, rtol=1e-5, atol=1e-5 * float(ref_logits.abs().max()))
    top2 = torch.topk(ref_logits, 2, dim=1)[0] if C > 1 else None
    safe = (top2[:, 0] - top2[:, 1]) > 1e-4 * ref_logits.abs().max()
    assert torch.equal(preds.cpu()[safe], ref_preds[safe])
    again = gate.predict(q)
    assert torch.equal(again[0], logits) and torch.equal(again[1], preds)        # deterministic


def test_our_predict_domain_mirror():
    """Same signature / return value as multimodal_evaluate.py:382."""
    from devqa_b200 import retrieval
    from oracle.st_shim import StubEncoder
    g = _g(3)
    enc = StubEncoder()
    W_rand = torch.randn(384, 2000, generator=g)
    Wo = torch.randn(2, 2000, generator=g)
    logits, preds = retrieval.our_predict_domain("what is on the table", "a cat", enc, Wo, W_rand, "cuda:0")
    q = torch.tensor(enc.encode("New Fact: what is on the table a cat\nPrompt: what is on the table a cat\n\n"))
    ref_logits, ref_preds = _ranpac_reference(q[None].double(), W_rand.double(), Wo.double())
    torch.testing.assert_close(logits.cpu().double(), ref_logits[0], rtol=1e-5, atol=1e-4)
    assert int(preds) == int(ref_preds[0]) and logits.shape == (2,)


def test_hice_probe_fuses_gate_memory_and_lookup():
    """devqa_bank_hice_probe == gate (torch) + run_classifier2's memory override + select_icl_samples' search."""
    from devqa_b200 import FactBank, ScopeGate
    L = devqa_b200._lib
    g = _g(11)
    d, M, n, k = 384, 5000, 11, 16
    W_rand = torch.randn(d, M, generator=g)
    Wo = torch.randn(2, M, generator=g) / M ** 0.5
    bank = torch.randn(20000, d, generator=g)
    mem = [torch.randn(1, d, generator=g) for _ in range(60)]
    q = torch.randn(n, d, generator=g)
    mq = torch.randn(n, d, generator=g)
    mq[3] = mem[7][0] + 0.05 * torch.randn(d, generator=g)              # probe 3 is a remembered m-loc sample
    fb = FactBank.from_embeddings(bank)
    gate = ScopeGate(Wo, W_rand, mloc_memory={'embeddings': mem}, threshold=0.81)
    logits, preds, mem_sim, scores, ids = gate.probe(fb, q, k, mem_query_embeddings=mq)
    ref_logits, ref_preds = _ranpac_reference(q.double(), W_rand.double(), Wo.double())
    memory_fea = torch.cat(mem, 0)
    memory_fea = memory_fea / torch.norm(memory_fea, p=2, dim=-1).unsqueeze(-1)
    sim = (mq / torch.norm(mq, p=2, dim=-1, keepdim=True)) @ memory_fea.T
    want = ref_preds.clone()
    want[sim.max(1)[0] > 0.81] = 0
    assert float(sim[3].max()) > 0.81 and int(preds[3]) == 0
    assert torch.equal(preds.cpu(), want)
    torch.testing.assert_close(mem_sim.cpu(), sim.max(1)[0], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(logits.cpu().double(), ref_logits, rtol=1e-5, atol=1e-5)
    s2, i2 = fb.search(q.cuda(), k, normalize_bank="noeps", normalize_query="noeps")
    assert torch.equal(ids, i2) and torch.equal(scores, s2)
    # no memory: NaN similarities, classifier decides alone
    gate2 = ScopeGate(Wo, W_rand)
    _, p2, ms2, _, _ = gate2.probe(fb, q, k)
    assert torch.isnan(ms2).all() and torch.equal(p2.cpu(), ref_preds)
    fb.close()


def test_hice_probe_batch_mirror():
    from devqa_b200 import retrieval
    from oracle.st_shim import StubEncoder
    import types
    g = _g(21)
    enc = StubEncoder()
    sents = [f"fact number {i} " + "x" * 300 for i in range(500)]
    emb = torch.tensor(enc.encode(sents))
    stored = {"sentences": sents, "embeddings": emb}
    W_rand = torch.randn(384, 3000, generator=g)
    Wo = torch.randn(2, 3000, generator=g)
    hp = types.SimpleNamespace(k=16, device=0)
    probes = [(f"fact number {i}", "x") for i in (3, 77, 123, 400)]
    got = retrieval.hice_probe_batch(probes, stored, enc, "cuda:0", hp, Wo, W_rand)
    for (prompt, target), (pred, icl) in zip(probes, got):
        _, ref_pred = retrieval.our_predict_domain(prompt, target, enc, Wo, W_rand, "cuda:0")
        assert pred == int(ref_pred)
        if pred != 0:
            assert icl == retrieval.select_icl_samples(prompt, target, stored, enc, "cuda:0", hp)
            assert len(''.join(icl)) <= 2500
        else:
            assert icl is None


def test_bank_grows_in_place_without_moving_rows():
    """regrow used to be cudaMalloc + a device-to-device copy of the whole bank.  With the bank in a
    reserved virtual range (csrc/vmm.cu) fresh physical chunks are mapped behind the stored rows: the data pointer stays
    put across many regrows, contents survive, searches before and after agree."""
    from devqa_b200 import FactBank
    g = _g(31)
    fb = FactBank(256, capacity=8)
    first = torch.randn(8, 256, generator=g)
    fb.append(first)
    ptr0 = fb.data_ptr
    ref = [first]
    moved = 0
    for step in range(40):
        new = torch.randn(1 + 997 * step, 256, generator=g)
        fb.append(new)
        ref.append(new)
        moved += int(fb.data_ptr != ptr0)
    whole = torch.cat(ref)
    assert len(fb) == whole.shape[0] > 700_000
    assert torch.equal(fb.rows(0, 8).cpu(), first) and torch.equal(fb.rows(len(fb) - 5).cpu(), whole[-5:])
    import os
    if os.environ.get("DEVQA_VMM", "1") != "0":
        assert moved == 0 and fb.data_ptr == ptr0, "the bank moved although virtual memory management is on"
    q = torch.randn(5, 256, generator=g)
    s, i = fb.search(q.cuda(), 8)
    s2, i2 = FactBank.from_embeddings(whole).search(q.cuda(), 8)
    assert torch.equal(i, i2) and torch.equal(s, s2)
    fb.truncate(100)
    fb.append(whole[:50])
    assert torch.equal(fb.rows(100, 150).cpu(), whole[:50])
    fb.close()
