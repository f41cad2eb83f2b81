"""A stand-in ``sentence_transformers`` module for running reference sources here.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The reference imports ``from sentence_transformers import SentenceTransformer, util``
(dataset/vllm.py:8, ike_main.py:1, multimodal_evaluate.py:394, HICE.py).  The
real package is absent from this image, so golden generation and the drop-in
tests install a module object of that name whose ``util`` is either the oracle
(golden generation) or the product shim (drop-in tests), and whose
``SentenceTransformer`` is a deterministic stub encoder: no MiniLM weights exist
on disk and the encoder is out of scope.
"""
from __future__ import annotations

import hashlib
import sys
import types
from typing import Iterable, Union

import numpy as np


class StubEncoder:
    """Deterministic text -> unit-norm float32 vector (sha256-seeded PCG64 normal).

    Mirrors the slice of the ``SentenceTransformer`` interface the reference
    touches: ``encode(str | list[str], show_progress_bar=, device=)`` returning
    ``np.float32`` (1-D for a str, 2-D for a list) and ``.to(device)``.
    all-MiniLM-L6-v2 emits 384-d unit-norm rows, hence the defaults.
    ``mix`` blends in a vector derived from the lower-cased word set so that
    textually close sentences are close in embedding space (gives the
    retrieval tests meaningful neighbours).
    """

    def __init__(self, name_or_path: str = "stub", device=None, dim: int = 384, mix: float = 0.7):
        self.name_or_path = name_or_path
        self.dim = dim
        self.mix = mix

    def to(self, device):
        return self

    def _vec(self, key: str) -> np.ndarray:
        seed = int.from_bytes(hashlib.sha256(key.encode("utf-8")).digest()[:8], "little")
        return np.random.Generator(np.random.PCG64(seed)).standard_normal(self.dim)

    def _encode_one(self, text: str) -> np.ndarray:
        v = (1.0 - self.mix) * self._vec("s:" + text)
        words = sorted(set(text.lower().split()))
        if words:
            w = sum(self._vec("w:" + t) for t in words) / np.sqrt(len(words))
            v = v + self.mix * w
        v = v / np.linalg.norm(v)
        return v.astype(np.float32)

    def encode(self, sentences: Union[str, Iterable[str]], show_progress_bar: bool = False,
               device=None, **_) -> np.ndarray:
        if isinstance(sentences, str):
            return self._encode_one(sentences)
        return np.stack([self._encode_one(s) for s in sentences]).astype(np.float32)


def install(util_module, encoder_cls=StubEncoder) -> types.ModuleType:
    """Register a fake ``sentence_transformers`` in ``sys.modules``.

    ``util_module`` must expose ``normalize_embeddings``, ``dot_score``,
    ``cos_sim`` and ``semantic_search``.
    """
    mod = types.ModuleType("sentence_transformers")
    mod.SentenceTransformer = encoder_cls
    mod.util = util_module
    mod.__dict__["__shim__"] = True
    sys.modules["sentence_transformers"] = mod
    sys.modules["sentence_transformers.util"] = util_module
    return mod


def uninstall() -> None:
    for name in ("sentence_transformers", "sentence_transformers.util"):
        m = sys.modules.get(name)
        if m is not None and (getattr(m, "__shim__", False) or name.endswith(".util")):
            sys.modules.pop(name, None)
