"""devqa_b200 -- B200-native fact-bank retrieval for DE-VQA / LiveEdit's IKE hot path.

The package holds only what that path needs:

  csrc/ + libdevqa.so   hand-written sm_100a CUDA kernels behind the C ABI of include/devqa.h
  _lib.py               ctypes binding of that ABI (fails loudly when the library is missing)
  bank.py               FactBank: resident, appendable bank + fused normalise/contract/top-k search
  util.py               drop-in for the slice of sentence_transformers.util the reference calls
  retrieval.py          host-side mirror of the reference's retrieval glue (finds_sim, IKE/HICE lookups)
  dist.py               row-sharded bank over the GPUs of one box (one all-gather + fused merge)
  pools.py              appended-pool top-1 / thresholded retrieval of the lifelong editors (LTE, SERAC, RECIPE),
                        LiveEdit's hard-routing similarity (retrieve_moes), find_closest_tokens over a resident vocabulary,
                        HICE's scope gate (RanPAC classifier + m-loc memory check) fused with the lookup
"""
from . import _lib
from .bank import FactBank, hits_from_arrays, norm_flags, topk_merge
from . import util
from . import retrieval
from . import dist
from .pools import EditPool, LiveEditPool, ScopeGate, VocabBank

__all__ = ["FactBank", "EditPool", "LiveEditPool", "ScopeGate", "VocabBank", "util", "retrieval", "dist", "hits_from_arrays", "norm_flags", "topk_merge", "_lib"]
