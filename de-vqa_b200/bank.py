"""Resident fact bank: the device-side replacement of the reference's per-request
``pickle.load -> torch.tensor(...).to(device) -> normalize`` bank handling and of its
``util.semantic_search`` call.

Reference call sites this object serves:
  * bank upload + normalise once  dataset/vllm.py:94-104,108-117
  * the same, but per request      easyeditor/models/ike/ike_main.py:31-37,187-194
  * whole-bank renormalise + search per probe  easyeditor/evaluate/multimodal_evaluate.py:406-412
  * ``torch.cat`` pool growth      editor/vllm_editors/lte_vl/lte_vl.py:150 (and serac/recipe)

All arithmetic happens in libdevqa.so (hand-written sm_100a CUDA) through the C ABI of
``include/devqa.h``; torch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os
import pickle
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib as L

NormMode = Union[bool, str, None]


def norm_flags(normalize_bank: NormMode, normalize_query: NormMode) -> int:
    """Translate ('eps' | True | 'noeps' | False | None) pairs into DEVQA_* flags.

    'eps'  = ``util.normalize_embeddings`` (F.normalize, clamp 1e-12; dataset/vllm.py:104)
    'noeps' = ``x / torch.norm(x)`` of the HICE call sites (ike_main.py:231,238)
    The C ABI has a single NOEPS bit, so mixed eps/noeps pairs are rejected.
    """
    def canon(m):
        if m in (None, False, "none"):
            return None
        if m in (True, "eps"):
            return "eps"
        if m == "noeps":
            return "noeps"
        raise ValueError(f"unknown normalisation mode {m!r}")
    b, q = canon(normalize_bank), canon(normalize_query)
    if b and q and b != q:
        raise ValueError("bank and query must use the same normalisation variant (eps / noeps)")
    flags = 0
    if b:
        flags |= L.NORMALIZE_BANK
    if q:
        flags |= L.NORMALIZE_QUERY
    if "noeps" in (b, q):
        flags |= L.NOEPS
    return flags


def _stream_ptr(device: int) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def hits_from_arrays(scores: np.ndarray, ids: np.ndarray) -> List[List[dict]]:
    """``[Q,k]`` arrays -> the reference's hit records:
    list[Q] of list[<=k] of ``{"corpus_id": int, "score": float}``, score-descending."""
    out = []
    for srow, irow in zip(scores.tolist(), ids.tolist()):
        out.append([{"corpus_id": i, "score": s} for s, i in zip(srow, irow) if i >= 0])
    return out


class FactBank:
    """An appendable ``[N, D]`` embedding bank resident in HBM, searched in place.

    ``dtype``: "fp32" (bit-for-bit the rows that were appended) or "bf16" (rows rounded to bf16
    on append; halves the bytes a search streams; scores still accumulate in fp32).
    """

    def __init__(self, dim: int, device: Union[int, str, torch.device] = 0, dtype: str = "fp32",
                 capacity: int = 0, shadow: bool = False):
        dev = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        if dev.type != "cuda":
            raise ValueError("FactBank lives on a CUDA device; there is no CPU path")
        self.device_index = dev.index if dev.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", self.device_index)
        self.dim = int(dim)
        self.dtype = {"fp32": L.F32, "float32": L.F32, "bf16": L.BF16, "bfloat16": L.BF16}[dtype]
        self._h = C.c_void_p()
        torch.cuda.init()
        L.check(L.lib().devqa_bank_create(C.byref(self._h), self.device_index, self.dim, int(capacity), self.dtype))
        if shadow:
            self.set_shadow(True)
        # text columns of the pickle layout (util.py:84-86); optional
        self.sentences: List[str] = []
        self.images: List[str] = []
        self.prompts: List[List[str]] = []

    # ---- construction from the reference's on-disk layouts --------------------------------------
    @classmethod
    def from_embeddings(cls, embeddings, device=0, dtype: str = "fp32", capacity: int = 0, shadow: bool = False) -> "FactBank":
        emb = embeddings if isinstance(embeddings, torch.Tensor) else torch.as_tensor(np.asarray(embeddings))
        if emb.dim() != 2:
            raise ValueError("embeddings must be [N, D]")
        bank = cls(emb.shape[1], device=device, dtype=dtype, capacity=max(capacity, emb.shape[0]), shadow=shadow)
        bank.append(emb)
        return bank

    def auto_shadow(self) -> bool:
        """Give a large fp32 bank its 2-byte shadow (``set_shadow``) if the scans that use it can occur (>= 148 tiles of
        128 rows), it fits in a quarter of the free device memory, and ``DEVQA_SHADOW`` is not ``0``.  Results do not
        change; searches of such banks stream half the bytes.  Returns whether the bank has a shadow afterwards."""
        if self.has_shadow:
            return True
        if self.dtype != L.F32 or len(self) < 148 * 128 or os.environ.get("DEVQA_SHADOW", "auto") == "0":
            return False
        need = len(self) * ((self.dim + 7) // 8 * 8) * 2 + len(self) * 4
        free, _total = torch.cuda.mem_get_info(self.device)
        if need > free // 4:
            return False
        try:
            self.set_shadow(True)
        except L.DevqaError:                                   # (out of memory after all: the bank works without one)
            return False
        return True

    @classmethod
    def from_stored_data(cls, stored: dict, device=0, dtype: str = "fp32", shadow="auto") -> "FactBank":
        """From the dict the reference pickles (util.py:27,50,85) or ``torch.save``s
        (multimodal_editor.py:1129-1130): keys ``sentences``, ``embeddings`` and optionally
        ``images``, ``prompts``.  ``shadow``: True / False / "auto" (``auto_shadow``)."""
        bank = cls.from_embeddings(stored["embeddings"], device=device, dtype=dtype, shadow=shadow is True)
        if shadow == "auto":
            bank.auto_shadow()
        bank.sentences = list(stored.get("sentences", []))
        bank.images = list(stored.get("images", []))
        bank.prompts = [list(p) for p in stored.get("prompts", [])]
        return bank

    @classmethod
    def from_pickle(cls, path: str, device=0, dtype: str = "fp32", shadow="auto") -> "FactBank":
        """Load ``{task}_embeddings(_llava).pkl`` / ``vlkeb_embeddings_collect.pkl`` unchanged.
        ``shadow``: True / False / "auto" (``auto_shadow``)."""
        if str(path).endswith((".pth", ".pt")):
            stored = torch.load(path, map_location="cpu", weights_only=False)
        else:
            with open(path, "rb") as f:
                stored = pickle.load(f)
        return cls.from_stored_data(stored, device=device, dtype=dtype, shadow=shadow)

    # ---- state -------------------------------------------------------------------------------
    def __len__(self) -> int:
        return int(L.lib().devqa_bank_size(self._h))

    @property
    def capacity(self) -> int:
        return int(L.lib().devqa_bank_capacity(self._h))

    @property
    def data_ptr(self) -> int:
        return int(L.lib().devqa_bank_data(self._h) or 0)

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            L.lib().devqa_bank_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_shadow(self, enable: bool = True) -> None:
        """Keep (or drop) a 2-byte (scaled fp16) shadow of this fp32 bank: large scans stream 2 bytes per element instead of 4 and
        re-score their candidates against the fp32 rows -- results stay the exact fp32 results, bit for bit."""
        L.check(L.lib().devqa_bank_set_shadow(self._h, 1 if enable else 0, _stream_ptr(self.device_index)))

    @property
    def has_shadow(self) -> bool:
        return bool(L.lib().devqa_bank_has_shadow(self._h))

    def set_id_map(self, id_base: int = 0, chunk_rows: int = 0, rank: int = 0, world: int = 1) -> None:
        L.check(L.lib().devqa_bank_set_id_map(self._h, int(id_base), int(chunk_rows), int(rank), int(world)))

    def reserve(self, rows: int) -> None:
        L.check(L.lib().devqa_bank_reserve(self._h, int(rows), _stream_ptr(self.device_index)))

    def truncate(self, rows: int) -> None:
        L.check(L.lib().devqa_bank_truncate(self._h, int(rows)))
        for col in (self.sentences, self.images, self.prompts):
            del col[rows:]

    def append(self, rows, sentences: Optional[Sequence[str]] = None, images: Optional[Sequence[str]] = None,
               prompts: Optional[Sequence[Sequence[str]]] = None) -> None:
        """Append ``[n, D]`` (or ``[D]``) fp32 rows in place -- amortised O(1) per row, no
        ``torch.cat`` of the whole pool (lte_vl.py:150).  Accepts CUDA tensors (device copy),
        CPU tensors and ndarrays (one H2D copy)."""
        t = rows if isinstance(rows, torch.Tensor) else torch.as_tensor(np.asarray(rows))
        if t.dim() == 1:
            t = t.unsqueeze(0)
        if t.dim() != 2 or t.shape[1] != self.dim:
            raise ValueError(f"rows must be [n, {self.dim}], got {tuple(t.shape)}")
        t = t.detach().to(torch.float32)
        if t.is_cuda and t.device.index != self.device_index:
            t = t.to(self.device)
        if t.stride(1) != 1 or (t.shape[0] > 1 and t.stride(0) < self.dim):
            t = t.contiguous()
        ld = t.stride(0) if t.shape[0] > 1 else self.dim
        L.check(L.lib().devqa_bank_append(self._h, t.data_ptr(), t.shape[0], ld, 0 if t.is_cuda else 1,
                                          _stream_ptr(self.device_index)))
        if sentences is not None:
            self.sentences.extend(sentences)
        if images is not None:
            self.images.extend(images)
        if prompts is not None:
            self.prompts.extend([list(p) for p in prompts])

    def append_synthetic(self, n_rows: int, seed: int, first_global_row: int = 0, kind: int = 0) -> None:
        L.check(L.lib().devqa_bank_append_synthetic(self._h, int(n_rows), int(seed), int(first_global_row), int(kind),
                                                    _stream_ptr(self.device_index)))

    def rows(self, start: int = 0, stop: Optional[int] = None) -> torch.Tensor:
        """Copy of rows [start, stop) as an fp32 CUDA tensor (tests / debugging)."""
        stop = len(self) if stop is None else stop
        n = max(0, stop - start)
        out = torch.empty((n, self.dim), dtype=torch.float32, device=self.device)
        L.check(L.lib().devqa_bank_read_rows(self._h, int(start), n, out.data_ptr(), self.dim,
                                             _stream_ptr(self.device_index)))
        return out

    # ---- search ------------------------------------------------------------------------------
    def _prep_queries(self, queries) -> torch.Tensor:
        q = queries if isinstance(queries, torch.Tensor) else torch.as_tensor(np.asarray(queries))
        if q.dim() == 1:
            q = q.unsqueeze(0)
        if q.dim() != 2 or q.shape[1] != self.dim:
            raise ValueError(f"queries must be [Q, {self.dim}], got {tuple(q.shape)}")
        q = q.detach().to(device=self.device, dtype=torch.float32)
        if q.stride(1) != 1 or (q.shape[0] > 1 and q.stride(0) < self.dim):
            q = q.contiguous()
        return q

    def search(self, queries, top_k: int, normalize_bank: NormMode = "eps", normalize_query: NormMode = "eps",
               flags: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
        """Tensor API (no host sync): ``(scores [Q,k] fp32, ids [Q,k] int64)`` on the bank's device.
        Slots beyond ``len(bank)`` hold ``-inf`` / ``-1``."""
        q = self._prep_queries(queries)
        nq = q.shape[0]
        scores = torch.empty((nq, top_k), dtype=torch.float32, device=self.device)
        ids = torch.empty((nq, top_k), dtype=torch.int64, device=self.device)
        ld = q.stride(0) if nq > 1 else self.dim
        L.check(L.lib().devqa_bank_search(self._h, q.data_ptr(), nq, ld, int(top_k),
                                          norm_flags(normalize_bank, normalize_query) | flags,
                                          scores.data_ptr(), ids.data_ptr(), _stream_ptr(self.device_index)))
        return scores, ids

    def search_host(self, queries: np.ndarray, top_k: int, normalize_bank: NormMode = "eps",
                    normalize_query: NormMode = "eps", flags: int = 0) -> Tuple[np.ndarray, np.ndarray]:
        """Host-buffer API: numpy in, numpy out; H2D, search, D2H and the sync happen inside the
        C call (``devqa_bank_search_host``)."""
        q = np.ascontiguousarray(np.asarray(queries, dtype=np.float32))
        if q.ndim == 1:
            q = q[None, :]
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ValueError(f"queries must be [Q, {self.dim}], got {q.shape}")
        scores = np.empty((q.shape[0], top_k), dtype=np.float32)
        ids = np.empty((q.shape[0], top_k), dtype=np.int64)
        L.check(L.lib().devqa_bank_search_host(self._h, q.ctypes.data, q.shape[0], self.dim, int(top_k),
                                               norm_flags(normalize_bank, normalize_query) | flags,
                                               scores.ctypes.data, ids.ctypes.data, _stream_ptr(self.device_index)))
        return scores, ids

    def hits(self, queries, top_k: int, normalize_bank: NormMode = "eps", normalize_query: NormMode = "eps"
             ) -> List[List[dict]]:
        """The reference's result format (what ``util.semantic_search`` returns)."""
        q = queries.detach().cpu().numpy() if isinstance(queries, torch.Tensor) else np.asarray(queries)
        scores, ids = self.search_host(q, top_k, normalize_bank, normalize_query)
        return hits_from_arrays(scores, ids)

    def search_packed(self, queries, top_k: int, normalize_bank: NormMode = "eps", normalize_query: NormMode = "eps",
                      flags: int = 0) -> torch.Tensor:
        """Local top-k as packed order-preserving keys ``[Q, k]`` (int64 view of uint64): the
        payload of the multi-GPU all-gather."""
        q = self._prep_queries(queries)
        nq = q.shape[0]
        keys = torch.empty((nq, top_k), dtype=torch.int64, device=self.device)
        ld = q.stride(0) if nq > 1 else self.dim
        L.check(L.lib().devqa_bank_search_packed(self._h, q.data_ptr(), nq, ld, int(top_k),
                                                 norm_flags(normalize_bank, normalize_query) | flags,
                                                 keys.data_ptr(), _stream_ptr(self.device_index)))
        return keys


def topk_merge(gathered: torch.Tensor, top_k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """``[G, Q, k]`` packed keys (each ``[k]`` run sorted descending) -> ``(scores, ids) [Q, top_k]``."""
    if gathered.dim() != 3 or gathered.dtype != torch.int64 or not gathered.is_cuda:
        raise ValueError("gathered must be a CUDA int64 tensor [G, Q, k]")
    gathered = gathered.contiguous()
    g, nq, k = gathered.shape
    if k != top_k:
        raise ValueError("list length must equal top_k")
    scores = torch.empty((nq, top_k), dtype=torch.float32, device=gathered.device)
    ids = torch.empty((nq, top_k), dtype=torch.int64, device=gathered.device)
    with torch.cuda.device(gathered.device):
        L.check(L.lib().devqa_topk_merge(gathered.data_ptr(), g, nq, int(top_k), scores.data_ptr(), ids.data_ptr(),
                                         _stream_ptr(gathered.device.index)))
    return scores, ids
