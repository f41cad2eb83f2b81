"""ctypes binding of libdevqa.so (include/devqa.h).  No fallback: a missing library raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdevqa.so")

# mirrors of include/devqa.h
OK = 0
E_INVALID, E_CUDA, E_NOMEM, E_UNSUPPORTED, E_NCCL, E_STATE = -1, -2, -3, -4, -5, -6
MAX_TOPK = 1024
F32, BF16 = 0, 1
NORMALIZE_BANK, NORMALIZE_QUERY, NOEPS = 1 << 0, 1 << 1, 1 << 2
FORCE_STREAM, FORCE_TENSOR, FORCE_TF32X3, FORCE_FILTER = 1 << 8, 1 << 9, 1 << 10, 1 << 11
NO_SHADOW = 1 << 15
DEBUG_TIGHT_BOUND = 1 << 12
DEBUG_NO_ONESHOT = 1 << 13
DEBUG_NO_EMIT = 1 << 3
DEBUG_FILTER_ONLY = 1 << 22

_p, _i, _i64, _u, _u64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint, C.c_uint64

# name -> (restype, argtypes); must list every symbol include/devqa.h declares
SIGNATURES = {
    "devqa_version": (_i, []),
    "devqa_last_error": (C.c_char_p, []),
    "devqa_launch_count": (_i64, []),
    "devqa_debug_tc_trace": (_i, [_p]),
    "devqa_filter_fallbacks": (_i, [_i, C.POINTER(_i64)]),
    "devqa_filter_bound_violations": (_i, [_i, C.POINTER(_i64)]),
    "devqa_filter_relerr": (C.c_double, [_i64, _i]),
    "devqa_profile_enable": (_i, [_i, _i]),
    "devqa_profile_collect": (_i, [_i, C.POINTER(C.c_double), C.POINTER(_i64), C.POINTER(_i64)]),
    "devqa_search_raw": (_i, [_p, _i, _i64, _i64, _i64, _p, _i64, _i64, _i, _u, _i64, _p, _p, _p]),
    "devqa_search_raw_ex": (_i, [_p, _i, _i64, _i64, _i64, _p, _i64, _i64, _i, _u, _i64, _p, _p, _p, _p, _p]),
    "devqa_search_raw_shadowed": (_i, [_p, _i, _i64, _i64, _i64, _p, _i64, _i64, _i, _u, _i64, _p, _p, _p, _i64, _p, _p, _p, _p]),
    "devqa_make_shadow": (_i, [_p, _i64, _i64, _i64, _p, _i64, _p, _p]),
    "devqa_row_sumsq": (_i, [_p, _i, _i64, _i64, _i64, _p, _p, _p]),
    "devqa_normalize_rows": (_i, [_p, _i64, _p, _i64, _i64, _i64, _u, _p]),
    "devqa_bank_create": (_i, [C.POINTER(_p), _i, _i64, _i64, _i]),
    "devqa_bank_destroy": (_i, [_p]),
    "devqa_bank_size": (_i64, [_p]),
    "devqa_bank_capacity": (_i64, [_p]),
    "devqa_bank_dim": (_i64, [_p]),
    "devqa_bank_dtype": (_i, [_p]),
    "devqa_bank_device": (_i, [_p]),
    "devqa_bank_data": (_p, [_p]),
    "devqa_bank_reserve": (_i, [_p, _i64, _p]),
    "devqa_bank_append": (_i, [_p, _p, _i64, _i64, _i, _p]),
    "devqa_bank_set_shadow": (_i, [_p, _i, _p]),
    "devqa_bank_has_shadow": (_i, [_p]),
    "devqa_bank_read_rows": (_i, [_p, _i64, _i64, _p, _i64, _p]),
    "devqa_bank_truncate": (_i, [_p, _i64]),
    "devqa_bank_set_id_map": (_i, [_p, _i64, _i64, _i, _i]),
    "devqa_bank_search": (_i, [_p, _p, _i64, _i64, _i, _u, _p, _p, _p]),
    "devqa_bank_search_host": (_i, [_p, _p, _i64, _i64, _i, _u, _p, _p, _p]),
    "devqa_bank_search_packed": (_i, [_p, _p, _i64, _i64, _i, _u, _p, _p]),
    "devqa_bank_select_above": (_i, [_p, _p, _i64, _i64, _p, C.c_float, _u, _p, _p, _p, _i64, _p, _p]),
    "devqa_scope_gate": (_i, [_p, _i64, _i64, _i64, _p, _i64, _i64, _p, _i64, _i, _u, _p, C.c_float, _p, _p, _p]),
    "devqa_bank_hice_probe": (_i, [_p, _p, _p, _p, _i64, _i64, _p, _i64, _i64, _p, _i64, _i, C.c_float, _i, _u,
                                   _p, _p, _p, _p, _p, _p]),
    "devqa_topk_merge": (_i, [_p, _i, _i64, _i, _p, _p, _p]),
    "devqa_dist_unique_id": (_i, [_p]),
    "devqa_dist_init": (_i, [C.POINTER(_p), _p, _i, _i, _i]),
    "devqa_dist_destroy": (_i, [_p]),
    "devqa_dist_p2p_export": (_i, [_p, _p]),
    "devqa_dist_p2p_connect": (_i, [_p, _p]),
    "devqa_dist_p2p_set": (_i, [_p, _i]),
    "devqa_dist_p2p_enabled": (_i, [_p]),
    "devqa_dist_p2p_timeouts": (_i, [_p, C.POINTER(_i64)]),
    "devqa_dist_rank": (_i, [_p]),
    "devqa_dist_world": (_i, [_p]),
    "devqa_bank_search_dist": (_i, [_p, _p, _p, _i64, _i64, _i, _u, _p, _p, _p]),
    "devqa_bank_search_dist_host": (_i, [_p, _p, _p, _i64, _i64, _i, _u, _p, _p, _p]),
    "devqa_fill_synthetic": (_i, [_p, _i64, _i64, _i64, _u64, _i64, _i, _p]),
    "devqa_bank_append_synthetic": (_i, [_p, _i64, _u64, _i64, _i, _p]),
}

_lib = None


class DevqaError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libdevqa error {code}: {message}")
        self.code = code


def lib() -> C.CDLL:
    """Load (once) and return the library; raise loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python de-vqa_b200/build.py` "
                "(there is no CPU or PyTorch fallback for the retrieval path)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(code: int) -> None:
    if code != OK:
        msg = lib().devqa_last_error().decode("utf-8", "replace")
        if code == E_INVALID:
            raise ValueError(f"libdevqa: {msg}")
        raise DevqaError(code, msg)
