"""torch-CPU restatement of the retrieval arithmetic of DE-VQA's hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PARITY UNPINNED for the
arithmetic: the algorithm lives in the third-party ``sentence_transformers.util``
(absent from /root/reference, version not pinned by the reference); this file
restates its published behaviour and is anchored on the
reference's call sites:

  * ``util.normalize_embeddings``   dataset/vllm.py:66,104,117; ike_main.py:37,41,194,199
  * manual ``x / ||x||`` (no eps)   ike_main.py:231,238; multimodal_evaluate.py:404,410;
                                    HICE.py:170,175
  * ``util.dot_score``              passed as ``score_function`` at every search site
  * ``util.semantic_search``        dataset/vllm.py:68; ike_main.py:44,202,240;
                                    multimodal_evaluate.py:412; HICE.py:177;
                                    easyeditor/dataset/vqa.py:176; coco_caption.py:174

Everything here runs on the CPU with plain torch / numpy.
"""
from __future__ import annotations

import heapq
from typing import Callable, List, Sequence, Union

import numpy as np
import torch

ArrayLike = Union[torch.Tensor, np.ndarray, Sequence]

QUERY_CHUNK = 100        # upstream default query_chunk_size
CORPUS_CHUNK = 500_000   # upstream default corpus_chunk_size


def _as_tensor(x: ArrayLike) -> torch.Tensor:
    """Input coercion the upstream helpers apply (ndarray / list -> tensor)."""
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, (np.ndarray, np.generic)):
        return torch.from_numpy(np.asarray(x))
    if isinstance(x, (list, tuple)) and len(x) and isinstance(x[0], torch.Tensor):
        return torch.stack(list(x))
    return torch.as_tensor(x)


def _as_matrix(x: ArrayLike) -> torch.Tensor:
    t = _as_tensor(x)
    return t.unsqueeze(0) if t.dim() == 1 else t


def normalize_embeddings(embeddings: torch.Tensor) -> torch.Tensor:
    """Row-wise ``x / max(||x||_2, 1e-12)`` -- ``F.normalize(p=2, dim=1)``.

    Reference use: dataset/vllm.py:104,117 (bank, once) and :66 (query).
    """
    return torch.nn.functional.normalize(embeddings, p=2, dim=1)


def normalize_noeps(embeddings: torch.Tensor) -> torch.Tensor:
    """The HICE-side variant: ``x / ||x||`` with no clamp (zero row -> NaN).

    Reference: ike_main.py:238, multimodal_evaluate.py:410, HICE.py:170.
    """
    return embeddings / torch.norm(embeddings, p=2, dim=-1).unsqueeze(-1)


def dot_score(a: ArrayLike, b: ArrayLike) -> torch.Tensor:
    """``[len(a), len(b)]`` matrix of plain dot products, fp32 ``torch.mm``."""
    return torch.mm(_as_matrix(a), _as_matrix(b).transpose(0, 1))


def cos_sim(a: ArrayLike, b: ArrayLike) -> torch.Tensor:
    """Upstream's default ``score_function``: normalise both sides, then mm."""
    return torch.mm(normalize_embeddings(_as_matrix(a)),
                    normalize_embeddings(_as_matrix(b)).transpose(0, 1))


def semantic_search(query_embeddings: ArrayLike,
                    corpus_embeddings: ArrayLike,
                    query_chunk_size: int = QUERY_CHUNK,
                    corpus_chunk_size: int = CORPUS_CHUNK,
                    top_k: int = 10,
                    score_function: Callable = cos_sim) -> List[List[dict]]:
    """Chunked exhaustive top-k search, result as list[Q] of list[<=k] of
    ``{"corpus_id": int, "score": float}``, score-descending.

    Control flow restated from the upstream helper:
    (query chunk x corpus chunk) tiles -> ``score_function`` -> unsorted
    ``torch.topk`` of ``min(top_k, n)`` -> Python lists -> per-query bounded
    min-heap of ``(score, corpus_id)`` tuples -> dicts -> stable sort by score
    descending.  Order among exactly equal scores is implementation-defined and
    must not be parity-tested.
    """
    queries = _as_matrix(query_embeddings)
    corpus = _as_matrix(corpus_embeddings)
    if corpus.device != queries.device:
        queries = queries.to(corpus.device)

    n_q, n_c = queries.shape[0], corpus.shape[0]
    heaps: List[list] = [[] for _ in range(n_q)]

    for q0 in range(0, n_q, query_chunk_size):
        q_tile = queries[q0:q0 + query_chunk_size]
        for c0 in range(0, n_c, corpus_chunk_size):
            scores = score_function(q_tile, corpus[c0:c0 + corpus_chunk_size])
            keep = min(top_k, scores.shape[1])
            vals, cols = torch.topk(scores, keep, dim=1, largest=True, sorted=False)
            vals, cols = vals.cpu().tolist(), cols.cpu().tolist()
            for r, (row_vals, row_cols) in enumerate(zip(vals, cols)):
                heap = heaps[q0 + r]
                for col, val in zip(row_cols, row_vals):
                    entry = (val, c0 + col)
                    if len(heap) < top_k:
                        heapq.heappush(heap, entry)
                    else:
                        heapq.heappushpop(heap, entry)

    out: List[List[dict]] = []
    for heap in heaps:
        hits = [{"corpus_id": cid, "score": sc} for sc, cid in heap]
        hits.sort(key=lambda h: h["score"], reverse=True)
        out.append(hits)
    return out


# ---------------------------------------------------------------------------
# Array-form helpers used by the parity tests.
# ---------------------------------------------------------------------------

def search_arrays(queries: ArrayLike, bank: ArrayLike, top_k: int,
                  normalize_bank: str = "none", normalize_query: str = "none"):
    """Run the reference composition and return ``(scores[Q,k'] f32, ids[Q,k'] i64)``.

    ``normalize_*``: "none" | "eps" (``normalize_embeddings``) | "noeps"
    (the HICE variant).  This is exactly what the reference call sites do:
    normalise (a1/a2) then ``semantic_search(..., score_function=dot_score)`` (a4).
    """
    q = _as_matrix(queries).float()
    b = _as_matrix(bank).float()
    norm = {"none": lambda t: t, "eps": normalize_embeddings, "noeps": normalize_noeps}
    q = norm[normalize_query](q)
    b = norm[normalize_bank](b)
    hits = semantic_search(q, b, score_function=dot_score, top_k=top_k)
    kk = min(top_k, b.shape[0])
    scores = np.zeros((q.shape[0], kk), dtype=np.float32)
    ids = np.zeros((q.shape[0], kk), dtype=np.int64)
    for i, row in enumerate(hits):
        for j, h in enumerate(row):
            scores[i, j] = h["score"]
            ids[i, j] = h["corpus_id"]
    return scores, ids


def truth_scores_fp64(queries: ArrayLike, bank: ArrayLike,
                      normalize_bank: bool = True, normalize_query: bool = True,
                      eps: float = 1e-12) -> np.ndarray:
    """fp64 "truth" score matrix ``[Q, N]`` used to measure score gaps.

    The north star conditions index equality on "the reference's score gap
    exceeds 1e-5"; fp64 scores give that gap without fp32 accumulation noise.
    """
    q = np.asarray(_as_matrix(queries).double().numpy())
    b = np.asarray(_as_matrix(bank).double().numpy())
    if normalize_query:
        q = q / np.maximum(np.linalg.norm(q, axis=1, keepdims=True), eps)
    if normalize_bank:
        b = b / np.maximum(np.linalg.norm(b, axis=1, keepdims=True), eps)
    return q @ b.T


def compare_topk(test_scores: np.ndarray, test_ids: np.ndarray,
                 truth: np.ndarray, top_k: int,
                 gap: float = 1e-5, rtol: float = 1e-5, atol: float = 1e-6):
    """Gap-conditioned parity check of one implementation against fp64 truth.

    For every query: positions whose fp64 score is separated from both ranked
    neighbours (and from the best excluded score) by more than ``gap`` must
    carry exactly the truth index; every returned score must equal the fp64
    score of the row it names within ``rtol`` relative (+ ``atol`` for scores
    near zero); the returned *set* may differ from the truth set only in rows
    whose truth score lies within ``gap`` of the k-th score.

    Returns a dict of counters; raises AssertionError on a violation.
    """
    n_q, n = truth.shape
    kk = min(top_k, n)
    assert test_ids.shape == (n_q, kk), (test_ids.shape, (n_q, kk))
    order = np.argsort(-truth, axis=1, kind="stable")[:, :min(kk + 1, n)]
    pinned = 0
    for i in range(n_q):
        t_ids = order[i, :kk]
        t_sc = truth[i, order[i]]
        got_ids = test_ids[i]
        # scores must belong to the rows they name
        ref_sc = truth[i, got_ids]
        err = np.abs(test_scores[i].astype(np.float64) - ref_sc)
        tol = rtol * np.abs(ref_sc) + atol
        assert np.all(err <= tol), (
            f"query {i}: score mismatch max err {err.max():.3e} (tol {tol[err.argmax()]:.3e})")
        # descending order
        assert np.all(np.diff(test_scores[i]) <= 0), f"query {i}: scores not descending"
        # no duplicates
        assert len(set(got_ids.tolist())) == kk, f"query {i}: duplicate ids"
        for j in range(kk):
            up = t_sc[j - 1] - t_sc[j] if j > 0 else np.inf
            dn = t_sc[j] - t_sc[j + 1] if j + 1 < len(t_sc) else np.inf
            if up > gap and dn > gap:
                pinned += 1
                assert got_ids[j] == t_ids[j], (
                    f"query {i} rank {j}: got id {got_ids[j]} want {t_ids[j]} "
                    f"(gaps {up:.3e}/{dn:.3e})")
        # set difference only inside the gap band around the k-th score
        kth = t_sc[kk - 1]
        for rid in set(got_ids.tolist()) ^ set(t_ids.tolist()):
            assert abs(truth[i, rid] - kth) <= gap, (
                f"query {i}: id {rid} differs from truth set outside the gap band")
    return {"queries": n_q, "k": kk, "pinned_positions": pinned}
