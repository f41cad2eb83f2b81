"""Drop-in for the slice of ``sentence_transformers.util`` the reference uses.

Same names, signatures and result format as the helpers DE-VQA calls
(SURVEY.md 8b and Appendix A), so that reference call sites run unmodified when this
module is provided as ``sentence_transformers.util``:

    hits = util.semantic_search(q, bank, score_function=util.dot_score, top_k=k)
        dataset/vllm.py:68; easyeditor/models/ike/ike_main.py:44,202,240;
        easyeditor/evaluate/multimodal_evaluate.py:412; easyeditor/trainer/algs/HICE.py:177

Everything is computed by libdevqa.so on the GPU.  There is no CPU path: CPU tensors /
ndarrays are uploaded, and a machine without a B200 raises.  Divergences from upstream,
all outside the parity contract: ties are ordered by ascending
corpus id; ``score_function`` must be ``dot_score`` or
``cos_sim`` (the only two the reference passes or defaults to); chunk-size arguments are
accepted and ignored (the scan never materialises a score matrix).
"""
from __future__ import annotations

import os
import weakref
from typing import Callable, List

import numpy as np
import torch

from . import _lib as L
from .bank import _stream_ptr, hits_from_arrays


def _to_tensor(x) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, (np.ndarray, np.generic)):
        return torch.from_numpy(np.asarray(x))
    if isinstance(x, (list, tuple)) and len(x) and isinstance(x[0], torch.Tensor):
        return torch.stack(list(x))
    return torch.as_tensor(x)


def _matrix(x) -> torch.Tensor:
    t = _to_tensor(x)
    return t.unsqueeze(0) if t.dim() == 1 else t


def _cuda_f32(t: torch.Tensor, device=None) -> torch.Tensor:
    if not torch.cuda.is_available():
        raise RuntimeError("devqa_b200 has no CPU fallback: a CUDA device (B200, sm_100a) is required")
    if device is None:
        device = t.device if t.is_cuda else torch.device("cuda", torch.cuda.current_device())
    t = t.detach().to(device=device, dtype=torch.float32)
    if t.stride(-1) != 1 or (t.shape[0] > 1 and t.stride(0) < t.shape[1]):
        t = t.contiguous()
    return t


def _ld(t: torch.Tensor) -> int:
    return t.stride(0) if t.shape[0] > 1 else t.shape[1]


# ---- side information about caller-owned banks ----------------------------------------------------------------------
# The TF32 filter needs ||row||^2 per row (and their maximum) to bound its error.  A FactBank caches them on append;
# for a caller-owned tensor they are computed once per (storage, view, version) by one extra pass over the bank and
# kept for as long as the storage lives.  torch bumps a tensor's version counter on every in-place op (views share
# it), which is what invalidates an entry; a storage that dies takes its entries with it (weakref).
_NORM_CACHE_MIN_BYTES = 32 << 20          # below this the whole search is launch-latency bound anyway
_norm_cache: dict = {}

# A 2-byte SHADOW of the corpus (csrc/synth.cu shadow_convert_kernel) rides in the same cache entry: large
# scans then stream half the bytes and re-score their candidates against the caller's fp32 rows -- results stay
# bit-identical.  It costs rows x dim x 2 bytes of device memory for as long as the corpus tensor lives, so it is only
# made for corpora of at least 148 tiles (the scans that would use it), only while it fits in a quarter of the free
# memory, and never with DEVQA_LEVEL0_SHADOW=0.
_SHADOW_MIN_ROWS = 148 * 128


def _shadow_wanted(corpus: torch.Tensor) -> bool:
    if os.environ.get("DEVQA_LEVEL0_SHADOW", "1") == "0" or corpus.shape[0] < _SHADOW_MIN_ROWS:
        return False
    need = corpus.shape[0] * ((corpus.shape[1] + 7) // 8 * 8) * 2 + corpus.shape[0] * 4
    free, _total = torch.cuda.mem_get_info(corpus.device)
    return need <= free // 4


def _cached_row_norms(corpus: torch.Tensor, with_shadow: bool = False):
    """(row_sumsq device pointer, max-bits device pointer[, shadow tensor or None, scale tensor or None]) for an fp32 CUDA
    corpus, or Nones."""
    none = (None, None, None, None) if with_shadow else (None, None)
    if corpus.dtype != torch.float32 or corpus.numel() * 4 < _NORM_CACHE_MIN_BYTES:
        return none
    storage = corpus.untyped_storage()
    sid = storage._cdata
    view = (corpus.data_ptr(), tuple(corpus.shape), tuple(corpus.stride()))
    entry = _norm_cache.get(sid)
    if entry is None or entry[0]() is not storage:
        def _drop(_ref, sid=sid):
            _norm_cache.pop(sid, None)
        try:
            entry = _norm_cache[sid] = (weakref.ref(storage, _drop), {})
        except TypeError:
            return none
    hit = entry[1].get(view)
    if hit is None or hit[0] != corpus._version:
        n = corpus.shape[0]
        side = torch.zeros(n + 2, dtype=torch.float32, device=corpus.device)     # [n] sums, then the max word
        L.check(L.lib().devqa_row_sumsq(corpus.data_ptr(), L.F32, n, corpus.shape[1], _ld(corpus), side.data_ptr(),
                                        side.data_ptr() + 4 * n, _stream_ptr(corpus.device.index)))
        if len(entry[1]) >= 8:
            entry[1].clear()
        hit = entry[1][view] = [corpus._version, side, None, None]
    if with_shadow and hit[2] is None and _shadow_wanted(corpus):
        n, ld = corpus.shape[0], (corpus.shape[1] + 7) // 8 * 8
        try:
            shadow = torch.empty((n, ld), dtype=torch.float16, device=corpus.device)
            scale = torch.empty(n, dtype=torch.float32, device=corpus.device)
        except torch.cuda.OutOfMemoryError:
            shadow = scale = None
        if shadow is not None:
            L.check(L.lib().devqa_make_shadow(corpus.data_ptr(), n, corpus.shape[1], _ld(corpus), shadow.data_ptr(), ld,
                                              scale.data_ptr(), _stream_ptr(corpus.device.index)))
            hit[2], hit[3] = shadow, scale
    side = hit[1]
    if with_shadow:
        return side.data_ptr(), side.data_ptr() + 4 * corpus.shape[0], hit[2], hit[3]
    return side.data_ptr(), side.data_ptr() + 4 * corpus.shape[0]


def normalize_embeddings(embeddings: torch.Tensor) -> torch.Tensor:
    """``x / max(||x||_2, 1e-12)`` per row (``F.normalize(p=2, dim=1)``), on the GPU.
    Result lives on the input's device (reference: dataset/vllm.py:66,104,117)."""
    src = _to_tensor(embeddings)
    if src.dim() != 2:
        raise ValueError("normalize_embeddings expects a 2-D tensor")
    x = _cuda_f32(src)
    out = torch.empty_like(x, memory_format=torch.contiguous_format)
    with torch.cuda.device(x.device):
        L.check(L.lib().devqa_normalize_rows(x.data_ptr(), _ld(x), out.data_ptr(), x.shape[1], x.shape[0], x.shape[1],
                                             0, _stream_ptr(x.device.index)))
    return out if src.is_cuda else out.to(src.device)


def dot_score(a, b) -> torch.Tensor:
    """Marker + implementation of the reference's ``score_function``.  ``semantic_search``
    recognises this function and fuses it; called directly it returns the dense ``[len(a),
    len(b)]`` matrix with torch (off the hot path: no reference call site does that)."""
    return torch.mm(_matrix(a), _matrix(b).transpose(0, 1))


def cos_sim(a, b) -> torch.Tensor:
    """Upstream's default ``score_function`` (normalise both sides)."""
    a, b = _matrix(a), _matrix(b)
    return torch.mm(torch.nn.functional.normalize(a, p=2, dim=1),
                    torch.nn.functional.normalize(b, p=2, dim=1).transpose(0, 1))


def semantic_search_arrays(query_embeddings, corpus_embeddings, top_k: int = 10,
                           score_function: Callable = cos_sim, flags: int = 0):
    """Tensor form of ``semantic_search``: ``(scores [Q,k'] fp32, ids [Q,k'] int64)`` CUDA
    tensors with ``k' = min(top_k, N)``; no host synchronisation."""
    if score_function is dot_score:
        nflags = 0
    elif score_function is cos_sim:
        nflags = L.NORMALIZE_BANK | L.NORMALIZE_QUERY
    else:
        raise ValueError("devqa_b200.util.semantic_search supports score_function=dot_score or cos_sim only")
    corpus = _matrix(corpus_embeddings)
    queries = _matrix(query_embeddings)
    if corpus.dim() != 2 or queries.dim() != 2 or corpus.shape[1] != queries.shape[1]:
        raise ValueError(f"shape mismatch: queries {tuple(queries.shape)} vs corpus {tuple(corpus.shape)}")
    given = corpus
    corpus = _cuda_f32(corpus)
    # norms are cached only for the caller's own resident tensor (an uploaded / re-laid-out temporary dies with this call)
    resident = given.is_cuda and corpus.data_ptr() == given.data_ptr()
    queries = _cuda_f32(queries, device=corpus.device)      # upstream moves the query to the corpus device
    n, nq = corpus.shape[0], queries.shape[0]
    if n == 0:
        raise RuntimeError("semantic_search on an empty corpus (upstream raises in torch.topk as well)")
    kk = min(int(top_k), n)
    scores = torch.empty((nq, kk), dtype=torch.float32, device=corpus.device)
    ids = torch.empty((nq, kk), dtype=torch.int64, device=corpus.device)
    with torch.cuda.device(corpus.device):
        # batches the tensor-core path serves: hand it ||row||^2 so that it can filter (see _cached_row_norms)
        # (and, for large corpora, a 2-byte shadow of the rows: half the bytes to stream, same results)
        ss_ptr, max_ptr, shadow, scale = (_cached_row_norms(corpus, with_shadow=True) if (resident and kk <= 64 and
                                          (nq > 2 or n >= 8 * _SHADOW_MIN_ROWS)) else (None, None, None, None))
        L.check(L.lib().devqa_search_raw_shadowed(corpus.data_ptr(), L.F32, n, corpus.shape[1], _ld(corpus),
                                                  queries.data_ptr(), nq, _ld(queries), kk, nflags | flags, 0,
                                                  ss_ptr, max_ptr, shadow.data_ptr() if shadow is not None else None,
                                                  shadow.shape[1] if shadow is not None else 0,
                                                  scale.data_ptr() if scale is not None else None,
                                                  scores.data_ptr(), ids.data_ptr(), _stream_ptr(corpus.device.index)))
    return scores, ids


def semantic_search(query_embeddings, corpus_embeddings, query_chunk_size: int = 100,
                    corpus_chunk_size: int = 500000, top_k: int = 10,
                    score_function: Callable = cos_sim) -> List[List[dict]]:
    """Same contract as upstream: list[Q] of list[min(top_k, N)] of
    ``{"corpus_id": int, "score": float}``, score-descending.  One D2H + sync, like the
    reference's ``.cpu().tolist()``."""
    scores, ids = semantic_search_arrays(query_embeddings, corpus_embeddings, top_k, score_function)
    return hits_from_arrays(scores.cpu().numpy(), ids.cpu().numpy())
