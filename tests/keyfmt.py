"""Python restatement of the packed (score, id) key format of csrc/common.cuh (tests only)."""
import numpy as np


def pack_keys(scores: np.ndarray, ids: np.ndarray) -> np.ndarray:
    b = np.ascontiguousarray(scores, dtype=np.float32).view(np.uint32).astype(np.uint64)
    nan = (b & 0x7fffffff) > 0x7f800000
    neg = (b & 0x80000000) != 0
    code = np.where(neg, (~b) & 0xffffffff, b | 0x80000000)
    code = np.where(nan, 0xffffffff, code).astype(np.uint64)
    key = (code << np.uint64(32)) | (np.uint64(0xffffffff) - ids.astype(np.uint64))
    return np.where(ids < 0, np.uint64(0), key)


def unpack_keys(keys: np.ndarray):
    keys = keys.astype(np.uint64)
    code = (keys >> np.uint64(32)).astype(np.uint32)
    ids = (np.uint64(0xffffffff) - (keys & np.uint64(0xffffffff))).astype(np.int64)
    bits = np.where(code & 0x80000000, code & 0x7fffffff, ~code).astype(np.uint32)
    scores = bits.view(np.float32).copy()
    scores[code == 0xffffffff] = np.nan
    empty = keys == 0
    scores[empty] = -np.inf
    ids[empty] = -1
    return scores, ids


def merge_keys(gathered: np.ndarray, top_k: int) -> np.ndarray:
    """[G, Q, k] -> [Q, top_k]: what devqa_topk_merge computes."""
    g, q, k = gathered.shape
    flat = np.transpose(gathered, (1, 0, 2)).reshape(q, g * k)
    return np.sort(flat, axis=1)[:, ::-1][:, :top_k]
