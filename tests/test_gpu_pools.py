"""Appended-pool retrieval of the lifelong editors (N1) against the reference's torch composition."""
import pytest
import torch
import torch.nn.functional as F

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
