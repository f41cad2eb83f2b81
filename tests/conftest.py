import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    """GPU tests never run silently on a machine without a GPU: they are skipped with a reason
    unless selected with -m gpu, in which case a missing device is an error, not a skip."""
    import torch
    if torch.cuda.is_available():
        return
    selected_gpu = "gpu" in (config.getoption("-m") or "") and "not gpu" not in (config.getoption("-m") or "")
    for item in items:
        if "gpu" in item.keywords and not selected_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN, "evqa_golden.json")) as f:
        g = json.load(f)
    g["embeddings"] = np.load(os.path.join(GOLDEN, "evqa_bank.npz"))["embeddings"]
    return g


@pytest.fixture(scope="session")
def built_lib():
    """Build libdevqa.so if it is stale (nvcc cross-compiles without a GPU)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("devqa_build", os.path.join(ROOT, "de-vqa_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()
