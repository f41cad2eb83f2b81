"""The C-ABI library builds, loads and exports exactly what include/devqa.h declares.
No compute calls: this runs on machines without a GPU."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "devqa.h")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"DEVQA_API\s+[\w\s\*]+?\b(devqa_\w+)\s*\(", src)))


def test_header_declares_something():
    syms = declared_symbols()
    assert len(syms) >= 25 and "devqa_bank_search" in syms and "devqa_search_raw" in syms


def test_library_exports_every_declared_symbol(built_lib):
    out = subprocess.run(["nm", "-D", "--defined-only", built_lib], capture_output=True, text=True, check=True).stdout
    exported = sorted({ln.split()[-1] for ln in out.splitlines() if " T devqa_" in ln})
    assert exported == declared_symbols()


def test_binding_covers_header_and_loads(built_lib):
    import devqa_b200
    L = devqa_b200._lib
    assert sorted(L.SIGNATURES) == declared_symbols()
    lib = L.lib()
    assert lib.devqa_version() == 100
    assert isinstance(lib.devqa_launch_count(), int)


def test_header_constants_match_binding(built_lib):
    import devqa_b200
    L = devqa_b200._lib
    src = open(HEADER).read()
    def const(name):
        m = re.search(rf"#define\s+{name}\s+\(?(1u\s*<<\s*\d+|-?\d+)\)?", src)
        v = m.group(1)
        return eval(v.replace("u", "")) if "<<" in v else int(v)
    for cname, pyname in [("DEVQA_MAX_TOPK", "MAX_TOPK"), ("DEVQA_F32", "F32"), ("DEVQA_BF16", "BF16"),
                          ("DEVQA_NORMALIZE_BANK", "NORMALIZE_BANK"), ("DEVQA_NORMALIZE_QUERY", "NORMALIZE_QUERY"),
                          ("DEVQA_NOEPS", "NOEPS"), ("DEVQA_E_INVALID", "E_INVALID"), ("DEVQA_E_UNSUPPORTED", "E_UNSUPPORTED"),
                          ("DEVQA_E_NCCL", "E_NCCL"), ("DEVQA_FORCE_STREAM", "FORCE_STREAM"), ("DEVQA_FORCE_TENSOR", "FORCE_TENSOR"),
                          ("DEVQA_FORCE_TF32X3", "FORCE_TF32X3"), ("DEVQA_FORCE_FILTER", "FORCE_FILTER"),
                          ("DEVQA_DEBUG_FILTER_ONLY", "DEBUG_FILTER_ONLY")]:
        assert const(cname) == getattr(L, pyname), cname


def test_argument_errors_need_no_gpu(built_lib):
    """Pure argument validation returns DEVQA_E_INVALID with a message before touching CUDA."""
    import devqa_b200
    L = devqa_b200._lib
    lib = L.lib()
    assert lib.devqa_bank_truncate(None, 0) == L.E_INVALID
    assert b"null" in lib.devqa_last_error()
    assert lib.devqa_bank_size(None) == L.E_INVALID
    assert lib.devqa_topk_merge(None, 0, 1, 1, None, None, None) == L.E_INVALID
    with pytest.raises(ValueError):
        L.check(lib.devqa_bank_set_id_map(None, 0, 0, 0, 1))


def test_header_compiles_as_plain_c(tmp_path):
    """The boundary is a C ABI: the header must be consumable by a C compiler."""
    c = tmp_path / "t.c"
    c.write_text('#include "devqa.h"\nint main(void){return DEVQA_VERSION==100?0:1;}\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(c),
                    "-o", str(tmp_path / "t.o")], check=True)


def test_no_oracle_import_in_product():
    """The product path must never route through the oracle."""
    pkg = os.path.join(ROOT, "de-vqa_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert "/root/reference" not in text, f
