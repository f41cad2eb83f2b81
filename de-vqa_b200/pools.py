"""Appended-pool retrieval of the lifelong editors over a resident FactBank.

The reference's retrieval-style editors keep one embedding per applied edit in a tensor that grows by
``torch.cat([pool, new], 0)`` -- an O(N) copy per edit -- and look it up with a dense similarity + max /
sort at every forward:

  LTE     editor/vllm_editors/lte_vl/lte_vl.py:94-110 (cosine top-1, ``v > sim_threshold``), append :150
  SERAC   editor/vllm_editors/serac_vl/serac_vl.py:68-71 + modules.py:52-58 (dot / sqrt(d), top-1, ``v >= 10``),
          append :104
  RECIPE  editor/vllm_editors/recipe_vl/recipe_vl.py:86-95 (dot / sqrt(d), top-``retr_top_k`` entries that beat
          the prototype in row 0), append :163-164

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
        L.check(L.lib().devqa_search_raw(self.bank.data_ptr, self.bank.dtype, 1, self.dim, bank_ld,
                                         q.data_ptr(), q.shape[0], q.stride(0) if q.shape[0] > 1 else self.dim, 1, 0, 0,
                                         proto.data_ptr(), proto_id.data_ptr(), _stream_ptr(self.bank.device_index)))
        proto = proto / (self.dim ** 0.5)
        mask = sims > proto
        ids = torch.masked_select(order, mask)
        return list(torch.split(ids, mask.sum(1).tolist())), (sims, order)
