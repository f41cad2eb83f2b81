"""Row-sharded fact bank over the GPUs of one box.

One process per GPU (``torch.distributed`` provides rank/world and the host-side exchange of
the NCCL unique id; nothing else).  The bank is split by rows, queries are replicated, every
GPU scans its shard and emits its local top-k as packed 64-bit keys, ONE ``ncclAllGather``
(issued by libdevqa on the search stream) makes every GPU hold ``[G, Q, k]`` keys and a fused
merge kernel reduces them -- the parallel analogue of the reference's sequential 500 000-row
corpus chunks + heap merge inside ``util.semantic_search``.  Keys are totally ordered, so the
G-GPU result equals the 1-GPU result bit for bit.

Two layouts:
  * contiguous  -- rank r holds rows [bounds[r], bounds[r+1]) of a static bank;
  * chunked     -- an appendable bank whose fixed-size chunks are dealt round-robin, so that
                   global id = insertion order (lifelong editing, config C5).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import numpy as np
import torch

from . import _lib as L
from .bank import FactBank, NormMode, _stream_ptr, norm_flags


# ---- partition arithmetic (pure Python; covered by the CPU tests) --------------------------------------
def shard_bounds(n_rows: int, world: int) -> List[int]:
    """Contiguous split: ``bounds[r] .. bounds[r+1]`` are rank r's rows; sizes differ by <= 1."""
    base, extra = divmod(n_rows, world)
    bounds = [0]
    for r in range(world):
        bounds.append(bounds[-1] + base + (1 if r < extra else 0))
    return bounds


def chunk_owner(global_row: int, chunk_rows: int, world: int) -> int:
    return (global_row // chunk_rows) % world


def global_to_local(global_row: int, chunk_rows: int, world: int) -> Tuple[int, int]:
    """(owner rank, local row) of a global row in the chunked layout."""
    chunk = global_row // chunk_rows
    return chunk % world, (chunk // world) * chunk_rows + global_row % chunk_rows


def local_to_global(local_row: int, chunk_rows: int, rank: int, world: int) -> int:
    """Inverse of :func:`global_to_local`; the same formula libdevqa applies (devqa_bank_set_id_map)."""
    return ((local_row // chunk_rows) * world + rank) * chunk_rows + local_row % chunk_rows


def split_append(first_global_row: int, n_rows: int, chunk_rows: int, rank: int, world: int) -> List[Tuple[int, int]]:
    """Which slices ``[(offset_in_batch, count), ...]`` of an appended batch of ``n_rows`` rows
    starting at global id ``first_global_row`` belong to ``rank`` (in order)."""
    out = []
    g, end = first_global_row, first_global_row + n_rows
    while g < end:
        chunk_end = min(end, (g // chunk_rows + 1) * chunk_rows)
        if chunk_owner(g, chunk_rows, world) == rank:
            out.append((g - first_global_row, chunk_end - g))
        g = chunk_end
    return out


# ---- communicator ------------------------------------------------------------------------------------------
class Communicator:
    """libdevqa's own NCCL communicator (``devqa_dist_init``).  The 128-byte unique id is
    created on rank 0 and shipped through ``torch.distributed`` (any backend)."""

    def __init__(self, device: Optional[int] = None, p2p: bool = True):
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.device = torch.cuda.current_device() if device is None else int(device)
        buf = C.create_string_buffer(128)
        if self.rank == 0:
            L.check(L.lib().devqa_dist_unique_id(buf))
        box = [bytes(buf.raw)]
        dist.broadcast_object_list(box, src=0)
        self._h = C.c_void_p()
        idbuf = C.create_string_buffer(box[0], 128)
        L.check(L.lib().devqa_dist_init(C.byref(self._h), idbuf, self.rank, self.world, self.device))
        self.p2p = self._connect_mailboxes(dist) if p2p and self.world > 1 else False

    def _connect_mailboxes(self, dist) -> bool:
        """Peer-to-peer exchange of the local top-k (devqa_dist_p2p_*): export this rank's mailbox, gather every
        rank's IPC handle over torch.distributed, map them.  Every rank ends up with the same setting: if any rank
        cannot map a peer (no NVLink/P2P, IPC blocked by the container), all stay on ncclAllGather."""
        import os
        if os.environ.get("DEVQA_P2P", "1") == "0" or self.world > 16:
            return False
        handle = C.create_string_buffer(64)
        ok = L.lib().devqa_dist_p2p_export(self._h, handle) == L.OK
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle.raw) if ok else None)
        if ok and all(h is not None for h in handles):
            ok = L.lib().devqa_dist_p2p_connect(self._h, C.create_string_buffer(b"".join(handles), 64 * self.world)) == L.OK
        else:
            ok = False
        verdict = [None] * self.world
        dist.all_gather_object(verdict, bool(ok))
        everyone = all(verdict)
        if ok and not everyone:
            L.lib().devqa_dist_p2p_set(self._h, 0)
        return everyone

    def p2p_timeouts(self) -> int:
        n = C.c_int64()
        L.check(L.lib().devqa_dist_p2p_timeouts(self._h, C.byref(n)))
        return int(n.value)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            L.lib().devqa_dist_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ShardedFactBank:
    """This rank's shard of a row-sharded bank + the collective search."""

    def __init__(self, dim: int, comm: Communicator, dtype: str = "fp32", chunk_rows: int = 0, capacity: int = 0,
                 shadow: bool = False):
        self.comm = comm
        self.rank, self.world = comm.rank, comm.world
        self.dim = dim
        self.chunk_rows = int(chunk_rows)
        self.local = FactBank(dim, device=comm.device, dtype=dtype, capacity=capacity, shadow=shadow)
        self.total_rows = 0
        if self.chunk_rows > 0:
            self.local.set_id_map(0, self.chunk_rows, self.rank, self.world)

    @classmethod
    def from_global_embeddings(cls, embeddings, comm: Communicator, dtype: str = "fp32") -> "ShardedFactBank":
        """Static bank, contiguous shards: every rank passes the same ``[N, D]`` array and keeps
        its own slice."""
        n = len(embeddings)
        bounds = shard_bounds(n, comm.world)
        lo, hi = bounds[comm.rank], bounds[comm.rank + 1]
        sb = cls(embeddings.shape[1], comm, dtype=dtype, capacity=hi - lo)
        sb.local.set_id_map(lo, 0, 0, 1)
        if hi > lo:
            sb.local.append(embeddings[lo:hi])
        sb.total_rows = n
        return sb

    @classmethod
    def synthetic(cls, n_rows: int, dim: int, comm: Communicator, seed: int, kind: int = 0, dtype: str = "fp32",
                  batch_rows: int = 1 << 20, shadow: bool = False) -> "ShardedFactBank":
        """Contiguous shards of the counter-based synthetic bank (any world size reproduces the
        same global bank)."""
        bounds = shard_bounds(n_rows, comm.world)
        lo, hi = bounds[comm.rank], bounds[comm.rank + 1]
        sb = cls(dim, comm, dtype=dtype, capacity=hi - lo, shadow=shadow)
        sb.local.set_id_map(lo, 0, 0, 1)
        for r0 in range(lo, hi, batch_rows):
            sb.local.append_synthetic(min(batch_rows, hi - r0), seed, first_global_row=r0, kind=kind)
        sb.total_rows = n_rows
        return sb

    def append(self, rows) -> None:
        """Chunked layout: every rank is handed the same batch; each keeps the chunks it owns."""
        if self.chunk_rows <= 0:
            raise RuntimeError("append needs the chunked layout (chunk_rows > 0)")
        t = rows if isinstance(rows, torch.Tensor) else torch.as_tensor(np.asarray(rows))
        if t.dim() == 1:
            t = t.unsqueeze(0)
        for off, cnt in split_append(self.total_rows, t.shape[0], self.chunk_rows, self.rank, self.world):
            self.local.append(t[off:off + cnt])
        self.total_rows += t.shape[0]

    def __len__(self) -> int:
        return self.total_rows

    def search(self, queries, top_k: int, normalize_bank: NormMode = "eps", normalize_query: NormMode = "eps",
               flags: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
        q = self.local._prep_queries(queries)
        nq = q.shape[0]
        scores = torch.empty((nq, top_k), dtype=torch.float32, device=self.local.device)
        ids = torch.empty((nq, top_k), dtype=torch.int64, device=self.local.device)
        ld = q.stride(0) if nq > 1 else self.dim
        L.check(L.lib().devqa_bank_search_dist(self.local._h, self.comm._h, q.data_ptr(), nq, ld, int(top_k),
                                               norm_flags(normalize_bank, normalize_query) | flags,
                                               scores.data_ptr(), ids.data_ptr(),
                                               _stream_ptr(self.local.device_index)))
        return scores, ids

    def search_host(self, queries: np.ndarray, top_k: int, normalize_bank: NormMode = "eps",
                    normalize_query: NormMode = "eps", flags: int = 0) -> Tuple[np.ndarray, np.ndarray]:
        q = np.ascontiguousarray(np.asarray(queries, dtype=np.float32))
        if q.ndim == 1:
            q = q[None, :]
        scores = np.empty((q.shape[0], top_k), dtype=np.float32)
        ids = np.empty((q.shape[0], top_k), dtype=np.int64)
        L.check(L.lib().devqa_bank_search_dist_host(self.local._h, self.comm._h, q.ctypes.data, q.shape[0], self.dim,
                                                    int(top_k), norm_flags(normalize_bank, normalize_query) | flags,
                                                    scores.ctypes.data, ids.ctypes.data,
                                                    _stream_ptr(self.local.device_index)))
        return scores, ids
