"""Appended-pool retrieval of the lifelong editors over a resident FactBank.

The reference's retrieval-style editors keep one embedding per applied edit in a tensor that grows by
``torch.cat([pool, new], 0)`` -- an O(N) copy per edit -- and look it up with a dense similarity + max /
sort at every forward:

  LTE     editor/vllm_editors/lte_vl/lte_vl.py:94-110 (cosine top-1, ``v > sim_threshold``), append :150
  SERAC   editor/vllm_editors/serac_vl/serac_vl.py:68-71 + modules.py:52-58 (dot / sqrt(d), top-1, ``v >= 10``),
          append :104
  RECIPE  editor/vllm_editors/recipe_vl/recipe_vl.py:86-95 (dot / sqrt(d), top-``retr_top_k`` entries that beat
          the prototype in row 0), append :163-164
  HICE    easyeditor/trainer/algs/HICE.py:238-250 (m-loc memory gate: no-eps cosine max > 0.81; part of N4)

``EditPool`` keeps the pool in a FactBank (in-place append, amortised O(1); ``restore`` = truncate) and serves
the three lookups with the same fused scan kernels as the IKE path.  Same decision rules as the reference;
similarities are fp32 within 1e-5 of the torch composition.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from . import _lib as L
from .bank import FactBank, _stream_ptr


class EditPool:
    def __init__(self, dim: int, device=0, capacity: int = 1024):
        self.bank = FactBank(dim, device=device, capacity=capacity)
        self.dim = dim

    def __len__(self) -> int:
        return len(self.bank)

    def append(self, reps) -> None:
        """``edit_one_piece``: add the new edit's representation(s) ``[n, d]`` in place."""
        self.bank.append(reps)

    def restore(self, n_rows: int = 0) -> None:
        """``restore_to_original_model``: drop the edits applied after the first ``n_rows``."""
        self.bank.truncate(n_rows)

    # ---- LTE ---------------------------------------------------------------------------------------
    def lte_retrieve(self, text_embds, sim_threshold: float) -> Tuple[List[Optional[int]], torch.Tensor]:
        """Cosine top-1 per query; index if ``sim > sim_threshold`` else None (lte_vl.py:97-108).
        Returns (indices, best similarity [b])."""
        if len(self.bank) == 0:
            q = torch.as_tensor(text_embds)
            return [None] * (1 if q.dim() == 1 else q.shape[0]), torch.empty(0)
        s, i = self.bank.search(text_embds, 1, normalize_bank="eps", normalize_query="eps")
        s, i = s[:, 0].cpu(), i[:, 0].cpu()
        return [int(ix) if float(v) > sim_threshold else None for v, ix in zip(s, i)], s

    # ---- HICE memory gate (N4) ------------------------------------------------------------------------
    def hice_memory_gate(self, query_embedding, threshold: float = 0.81) -> Tuple[bool, float]:
        """``HICE.run_classifier2`` (easyeditor/trainer/algs/HICE.py:238-250): ``sim = q/||q|| @ (mem/||mem||).T``
        over the m-loc memory, in scope iff ``sim.max() > threshold``.  The reference re-concatenates and
        re-normalises the whole memory on every call; here it is a resident pool (``append`` per remembered
        embedding) and the gate is one fused no-eps cosine top-1.  Returns (gate, max similarity)."""
        if len(self.bank) == 0:
            return False, float("-inf")
        s, _ = self.bank.search(query_embedding, 1, normalize_bank="noeps", normalize_query="noeps")
        best = float(s.max())
        return best > threshold, best

    # ---- SERAC -------------------------------------------------------------------------------------
    def serac_retrieve(self, rep1) -> Tuple[torch.Tensor, torch.Tensor]:
        """``sim = rep1 @ pool.T / sqrt(d)``; ``torch.max(sim, 1)`` -> (values [b], indices [b]) on the device
        (serac_vl.py:70-71, modules.py:52-58).  The caller applies its ``v >= 10`` gate."""
        s, i = self.bank.search(rep1, 1, normalize_bank=None, normalize_query=None)
        return s[:, 0] / (self.dim ** 0.5), i[:, 0]

    # ---- RECIPE ------------------------------------------------------------------------------------
    def recipe_retrieve(self, query_reps, retr_top_k: int) -> Tuple[List[torch.Tensor], Tuple[torch.Tensor, torch.Tensor]]:
        """Top-``retr_top_k`` pool rows whose scaled similarity exceeds the similarity with the prototype
        stored in row 0 (recipe_vl.py:87-92).  Returns (list of id tensors per query, (sorted_sim, order)
        truncated to ``retr_top_k`` columns)."""
        k = min(int(retr_top_k), len(self.bank))
        sims, order = self.bank.search(query_reps, k, normalize_bank=None, normalize_query=None)
        sims = sims / (self.dim ** 0.5)
        q = self.bank._prep_queries(query_reps)
        proto = torch.empty((q.shape[0], 1), dtype=torch.float32, device=self.bank.device)
        proto_id = torch.empty((q.shape[0], 1), dtype=torch.int64, device=self.bank.device)
        bank_ld = (self.dim + 3) // 4 * 4                      # FactBank pads fp32 rows to 16 bytes
        # similarity with the prototype = a one-row search over row 0 of the same resident buffer
        # (devqa_search_raw works on the CURRENT device: make that the bank's)
        with torch.cuda.device(self.bank.device):
            L.check(L.lib().devqa_search_raw(self.bank.data_ptr, self.bank.dtype, 1, self.dim, bank_ld,
                                             q.data_ptr(), q.shape[0], q.stride(0) if q.shape[0] > 1 else self.dim, 1, 0, 0,
                                             proto.data_ptr(), proto_id.data_ptr(), _stream_ptr(self.bank.device_index)))
        proto = proto / (self.dim ** 0.5)
        mask = sims > proto
        ids = torch.masked_select(order, mask)
        return list(torch.split(ids, mask.sum(1).tolist())), (sims, order)


class ScopeGate:
    """N4: HICE's scope decision for the probes of one edit request, resident on the device.

    Reference: ``our_predict_domain`` (easyeditor/evaluate/multimodal_evaluate.py:382-392) --
    ``relu(q/||q|| @ W_rand) @ Wo.T`` + argmax -- and the m-loc memory check of ``HICE.run_classifier2``
    (easyeditor/trainer/algs/HICE.py:238-250).  ``W_rand`` ``[dim, M]`` and ``Wo`` ``[n_classes, M]`` are uploaded once
    (the reference re-sends them per call as arguments of a Python function); ``mloc_memory`` (a list of ``[1, dim]``
    tensors, a ``[n, dim]`` tensor, or the reference's ``{'embeddings': [...]}`` dict) becomes a resident pool.
    ``predict`` is the gate alone; ``probe`` is gate + memory check + in-context lookup for ALL probes of a request in
    one C call (``devqa_bank_hice_probe``), replacing up to ~11 Python round trips of gate -> ``select_icl_samples``."""

    def __init__(self, Wo, W_rand, device=0, mloc_memory=None, threshold: float = 0.81):
        self.device = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        self.W_rand = torch.as_tensor(W_rand).detach().to(device=self.device, dtype=torch.float32).contiguous()
        self.Wo = torch.as_tensor(Wo).detach().to(device=self.device, dtype=torch.float32).contiguous()
        if self.W_rand.dim() != 2 or self.Wo.dim() != 2 or self.Wo.shape[1] != self.W_rand.shape[1]:
            raise ValueError("W_rand must be [dim, M] and Wo [n_classes, M]")
        self.dim, self.m = self.W_rand.shape
        self.n_classes = self.Wo.shape[0]
        self.threshold = float(threshold)
        self.memory: Optional[EditPool] = None
        if mloc_memory is not None:
            self.set_memory(mloc_memory)

    def set_memory(self, mloc_memory) -> None:
        if isinstance(mloc_memory, dict):
            mloc_memory = mloc_memory['embeddings']
        if isinstance(mloc_memory, (list, tuple)):
            mloc_memory = torch.cat([torch.as_tensor(e).reshape(-1, self.dim) for e in mloc_memory], 0) \
                if len(mloc_memory) else torch.zeros(0, self.dim)
        mem = torch.as_tensor(mloc_memory).reshape(-1, self.dim)
        self.memory = EditPool(self.dim, device=self.device, capacity=max(1, len(mem)))
        if len(mem):
            self.memory.append(mem)

    def remember(self, embedding) -> None:
        """``mloc_memory['embeddings'].append(...)``: one more remembered m-loc sample, in place."""
        if self.memory is None:
            self.memory = EditPool(self.dim, device=self.device)
        self.memory.append(torch.as_tensor(embedding).reshape(-1, self.dim))

    def _q(self, x) -> torch.Tensor:
        q = torch.as_tensor(x).detach().to(device=self.device, dtype=torch.float32)
        q = q.reshape(-1, self.dim)
        return q if q.is_contiguous() else q.contiguous()

    def predict(self, query_embeddings) -> Tuple[torch.Tensor, torch.Tensor]:
        """``our_predict_domain``'s arithmetic for ``[n, dim]`` (or ``[dim]``) probe embeddings:
        ``(logits [n, n_classes] fp32, preds [n] int64)`` on the device.  No memory check (the reference's
        ``our_predict_domain`` accepts ``mloc_memory`` but never reads it)."""
        q = self._q(query_embeddings)
        n = q.shape[0]
        logits = torch.empty((n, self.n_classes), dtype=torch.float32, device=self.device)
        preds = torch.empty((n,), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            L.check(L.lib().devqa_scope_gate(q.data_ptr(), n, self.dim, self.dim, self.W_rand.data_ptr(), self.m, self.m,
                                             self.Wo.data_ptr(), self.m, self.n_classes, L.NORMALIZE_QUERY | L.NOEPS,
                                             None, 0.0, logits.data_ptr(), preds.data_ptr(),
                                             _stream_ptr(self.device.index)))
        return logits, preds.long()

    def probe(self, bank: FactBank, query_embeddings, top_k: int, mem_query_embeddings=None,
              normalize_bank="noeps", normalize_query="noeps"):
        """All probes of a request at once: memory check (``run_classifier2``: the probe's "Prompt: ..." embedding
        ``mem_query_embeddings``, default the gate's own) -> scope gate -> top-k lookup over ``bank``
        (``select_icl_samples``'s no-eps cosine by default).  Returns device tensors
        ``(logits [n, C], preds [n] int64, mem_sim [n] (NaN without memory), scores [n, k], ids [n, k])``; the lookup
        of a probe matters only where ``preds != 0`` (multimodal_evaluate.py:498-505)."""
        from .bank import norm_flags
        q = self._q(query_embeddings)
        n = q.shape[0]
        mq = self._q(mem_query_embeddings) if mem_query_embeddings is not None else None
        if mq is not None and mq.shape != q.shape:
            raise ValueError("mem_query_embeddings must match query_embeddings in shape")
        if bank.device != self.device or bank.dim != self.dim:
            raise ValueError("bank lives on another device or has another width")
        logits = torch.empty((n, self.n_classes), dtype=torch.float32, device=self.device)
        preds = torch.empty((n,), dtype=torch.int32, device=self.device)
        mem_sim = torch.empty((n,), dtype=torch.float32, device=self.device)
        scores = torch.empty((n, top_k), dtype=torch.float32, device=self.device)
        ids = torch.empty((n, top_k), dtype=torch.int64, device=self.device)
        mem_h = self.memory.bank._h if self.memory is not None else None
        L.check(L.lib().devqa_bank_hice_probe(bank._h, mem_h, q.data_ptr(), mq.data_ptr() if mq is not None else None,
                                              n, self.dim, self.W_rand.data_ptr(), self.m, self.m, self.Wo.data_ptr(),
                                              self.m, self.n_classes, self.threshold, int(top_k),
                                              norm_flags(normalize_bank, normalize_query), logits.data_ptr(),
                                              preds.data_ptr(), mem_sim.data_ptr(), scores.data_ptr(), ids.data_ptr(),
                                              _stream_ptr(self.device.index)))
        return logits, preds.long(), mem_sim, scores, ids


class VocabBank:
    """N2: ``BaseVLLMForEdit.find_closest_tokens`` (editor/vllms_for_edit/base.py:187-196) --
    cosine top-k of hidden states against the vocabulary embedding matrix ``[V, d]`` -- with the matrix
    resident in a FactBank instead of being re-normalised (a full ``[V, d]`` temporary) on every call.

    The reference normalises the *queries* over ``dim=1`` of a ``[b, l, d]`` tensor, i.e. across the sequence
    axis, not the feature axis; parity means reproducing exactly that (it is a tiny ``[b, l, d]`` pre-scale
    done in torch), while the matrix rows get the fused ``x / ||x||`` (no eps) of the scan kernels."""

    def __init__(self, embedding_matrix, device=0, dtype: str = "fp32"):
        m = torch.as_tensor(embedding_matrix)
        self.bank = FactBank.from_embeddings(m.detach().to(torch.float32), device=device, dtype=dtype)
        self.dim = m.shape[1]

    def __len__(self) -> int:
        return len(self.bank)

    def close(self) -> None:
        self.bank.close()

    def find_closest_tokens(self, embeddings: torch.Tensor, top_k: int = 1) -> Tuple[torch.Tensor, torch.Tensor]:
        """embeddings ``[b, l, d]`` -> ``(toks [b, l, top_k] int64, sim [b, l, top_k] fp32)`` on the bank's device,
        ordered like ``torch.topk`` (descending)."""
        if embeddings.dim() != 3 or embeddings.shape[2] != self.dim:
            raise ValueError(f"embeddings must be [b, l, {self.dim}]")
        e = embeddings.detach().to(device=self.bank.device, dtype=torch.float32)
        e = e / e.norm(dim=1, keepdim=True)                     # base.py:190 (sic: the sequence axis)
        b, l, d = e.shape
        s, i = self.bank.search(e.reshape(b * l, d), int(top_k), normalize_bank="noeps", normalize_query=None)
        return i.view(b, l, top_k), s.view(b, l, top_k)


class LiveEditPool:
    """N3: the hard-routing similarity of LiveEdit's ``retrieve_moes``
    (editor/vllm_editors/liveedit/liveedit.py:132-153): ``einsum('bed,med->bme', ivr, evr_pool).mean(2) *
    sim_scale`` against a pool ``[m, eqe_n, module_dim]`` that grows by ``torch.cat`` per edit (:223-226),
    compared with the similarity to a prototype; every pool entry ABOVE it is routed to (a boolean mask, no
    top-k).  The mean over the ``eqe_n`` slots of per-slot dot products is one dot product over the flattened
    ``eqe_n * module_dim`` vector divided by ``eqe_n``, so the pool is a FactBank of that width, appended in
    place, and the mask comes out of the same streaming scan (all rows above the threshold, ascending ids --
    the order ``pool[retrieval_map]`` yields)."""

    def __init__(self, eqe_n: int, module_dim: int, device=0, capacity: int = 1024):
        self.eqe_n, self.module_dim = int(eqe_n), int(module_dim)
        self.dim = self.eqe_n * self.module_dim
        self.bank = FactBank(self.dim, device=device, capacity=capacity)

    def __len__(self) -> int:
        return len(self.bank)

    def append(self, evr) -> None:
        """``edit_one_piece`` (:223-226): ``evr`` is ``[n, eqe_n, module_dim]``."""
        t = torch.as_tensor(evr)
        self.bank.append(t.reshape(-1, self.dim))

    def restore(self, n_rows: int = 0) -> None:
        """``restore_to_original_model`` (:209-214)."""
        self.bank.truncate(n_rows)

    def similarities(self, ivr, sim_scale: float, top: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """The ``top`` largest ``vis_sim`` values and their pool ids (descending), device tensors."""
        q = torch.as_tensor(ivr).reshape(1, self.dim)
        s, i = self.bank.search(q, top, normalize_bank=None, normalize_query=None)
        return s[0] * (float(sim_scale) / self.eqe_n), i[0]

    def select(self, ivr, vis_sim_prot, sim_scale: float, cap: Optional[int] = None, want_mask: bool = True):
        """One ``devqa_bank_select_above`` call -- the bank-streaming kernel's threshold-select epilogue plus the
        ordered compaction (csrc/select.cu).  No host synchronisation: ``vis_sim_prot`` may be a device tensor (the
        prototype similarity the caller computed on the device, liveedit.py:140-141).  Returns device tensors
        ``(mask [m] bool | None, ids [cap] int64 ascending / -1, sims [cap] fp32, count [1] int64)``."""
        m = len(self.bank)
        dev = self.bank.device
        cap = m if cap is None else int(cap)
        q = self.bank._prep_queries(torch.as_tensor(ivr).reshape(1, self.dim))
        thr = torch.as_tensor(vis_sim_prot, dtype=torch.float32).reshape(-1)[:1].to(dev).contiguous()
        mask = torch.empty((1, m), dtype=torch.uint8, device=dev) if want_mask else None
        ids = torch.empty((1, max(cap, 1)), dtype=torch.int64, device=dev)
        sims = torch.empty((1, max(cap, 1)), dtype=torch.float32, device=dev)
        count = torch.empty((1,), dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            L.check(L.lib().devqa_bank_select_above(self.bank._h, q.data_ptr(), 1, self.dim, thr.data_ptr(),
                                                    float(sim_scale) / self.eqe_n, 0,
                                                    mask.data_ptr() if mask is not None and m > 0 else None,
                                                    ids.data_ptr(), sims.data_ptr(), cap, count.data_ptr(),
                                                    _stream_ptr(self.bank.device_index)))
        return (mask[0].bool() if mask is not None else None), ids[0, :cap], sims[0, :cap], count

    def retrieval_map(self, ivr, vis_sim_prot, sim_scale: float) -> torch.Tensor:
        """``retrieval_map = vis_sim > vis_sim_prot`` (liveedit.py:142) as a ``[m]`` bool device tensor; no host sync."""
        return self.select(ivr, vis_sim_prot, sim_scale, cap=0)[0]

    def retrieve(self, ivr, vis_sim_prot, sim_scale: float) -> Tuple[torch.Tensor, torch.Tensor]:
        """The ascending id list of the selected pool entries (the order ``pool[retrieval_map]`` yields) plus their
        similarities.  ``ivr``: ``[1, eqe_n, module_dim]``; ``vis_sim_prot``: the ``[1, 1]`` prototype similarity the
        caller computed (:140-141).  Any number of entries may pass (no top-k cap); the one host sync is reading the
        count to size the result, which ``pool[retrieval_map]`` also pays."""
        m = len(self.bank)
        dev = self.bank.device
        if m == 0:
            return torch.empty(0, dtype=torch.int64, device=dev), torch.empty(0, device=dev)
        _, ids, sims, count = self.select(ivr, vis_sim_prot, sim_scale, want_mask=False)
        n = int(count.item())
        return ids[:n], sims[:n]
