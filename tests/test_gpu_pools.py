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
