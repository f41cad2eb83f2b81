"""Parity of the tcgen05 / TMEM / TMA contraction kernel (large query batches) against the oracle,
and against the CUDA-core streaming kernel.  GPU only."""
import numpy as np
import pytest
import torch

import devqa_b200
from devqa_b200 import FactBank
from oracle import semantic_search as ss

pytestmark = pytest.mark.gpu
L = devqa_b200._lib
GAP, RTOL, ATOL = 1e-5, 1e-5, 1e-6


def _randn(n, d, seed, unit=True):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, d, generator=g)
    return torch.nn.functional.normalize(x, dim=1) if unit else x


def _hidden_like(n, d, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, d, generator=g) * torch.exp(0.5 * torch.randn(n, 1, generator=g))
    x[:, [1, d // 3, d // 2, d - 2]] *= 50.0
    return x


def _check_tc(bank_cpu, q_cpu, k, nb="eps", nq="eps"):
    """Both arithmetic variants of the tcgen05 scan: the default (TF32 filter + exact re-scoring, with the
    3xTF32 scan as its fallback) and the 3xTF32 scan alone."""
    fb = FactBank.from_embeddings(bank_cpu)
    kk = min(k, len(bank_cpu))
    truth = ss.truth_scores_fp64(q_cpu, bank_cpu, normalize_bank=bool(nb), normalize_query=bool(nq))
    for variant in (L.FORCE_TENSOR | L.FORCE_FILTER, L.FORCE_TENSOR | L.FORCE_TF32X3):
        s, i = fb.search(q_cpu.cuda(), k, normalize_bank=nb, normalize_query=nq, flags=variant)
        s, i = s.cpu().numpy(), i.cpu().numpy()
        stats = ss.compare_topk(s[:, :kk], i[:, :kk], truth, k, gap=GAP, rtol=RTOL, atol=ATOL)
        assert np.all(i[:, kk:] == -1)
    # the two kernels agree with each other under the same rule
    s2, i2 = fb.search(q_cpu.cuda(), k, normalize_bank=nb, normalize_query=nq, flags=L.FORCE_STREAM)
    ss.compare_topk(s2.cpu().numpy()[:, :kk], i2.cpu().numpy()[:, :kk], truth, k, gap=GAP, rtol=RTOL, atol=ATOL)
    fb.close()
    return stats


def test_tc_smallest_case():
    _check_tc(_randn(128, 32, 1), _randn(16, 32, 2), 4)


@pytest.mark.parametrize("n", [1, 5, 127, 128, 129, 300, 4097])
def test_tc_ragged_rows(n):
    _check_tc(_randn(n, 64, 10 + n), _randn(9, 64, 20 + n), 8)


@pytest.mark.parametrize("d", [4, 32, 36, 100, 384, 768, 1000, 4096])
def test_tc_dims(d):
    _check_tc(_randn(3000, d, 30 + d, unit=False), _randn(20, d, 40 + d, unit=False), 16)


@pytest.mark.parametrize("nq", [1, 9, 16, 17, 33, 64, 65, 100, 256])
def test_tc_query_counts(nq):
    _check_tc(_randn(5000, 256, 50 + nq), _randn(nq, 256, 60 + nq), 32)


@pytest.mark.parametrize("k", [1, 5, 16, 32, 33, 64, 100, 128])
def test_tc_topk_range(k):
    _check_tc(_randn(4000, 128, 70 + k), _randn(24, 128, 80 + k), k)


def test_tc_c2_shape_hidden_states():
    """C2: 8174 x 4096 un-normalised hidden-state-like rows, Q = 256, k = 16, fused normalise."""
    stats = _check_tc(_hidden_like(8174, 4096, 2234), _hidden_like(256, 4096, 6678), 16)
    assert stats["queries"] == 256


def test_tc_c3_shape():
    _check_tc(_randn(100_003, 768, 3234), _randn(64, 768, 7678), 32)


def test_tc_variants_and_edge_values():
    bank = _randn(2000, 96, 91, unit=False)
    bank[7] = 0
    bank[8] *= 1e4
    bank[9] *= 1e-4
    q = _randn(12, 96, 92, unit=False)
    _check_tc(bank, q, 10)
    _check_tc(bank, q, 10, nb=None, nq=None)
    _check_tc(bank, q, 10, nb="eps", nq=None)
    fb = FactBank.from_embeddings(bank)
    s, i = fb.search(q.cuda(), 4, normalize_bank="noeps", normalize_query="noeps", flags=L.FORCE_TENSOR)
    assert (i[:, 0] == 7).all() and torch.isnan(s[:, 0]).all()        # zero row -> NaN ranks first, as in the stream kernel
    fb.close()


def test_tc_is_deterministic_and_batch_independent():
    bank, q = _randn(30_000, 384, 101), _randn(40, 384, 102)
    fb = FactBank.from_embeddings(bank)
    a = fb.search(q.cuda(), 16, flags=L.FORCE_TENSOR)
    b = fb.search(q.cuda(), 16, flags=L.FORCE_TENSOR)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    c = fb.search(q[:20].cuda(), 16, flags=L.FORCE_TENSOR)
    assert torch.equal(c[1], a[1][:20]) and torch.equal(c[0], a[0][:20])
    fb.close()


def test_tc_unsupported_configs_fail_loudly():
    fb = FactBank.from_embeddings(_randn(100, 64, 1))
    with pytest.raises(L.DevqaError):
        fb.search(_randn(2, 64, 2).cuda(), 200, flags=L.FORCE_TENSOR)      # top_k capacity > 128
    fb.close()


# ---- bf16-bank variant (kind::f16, queries split three ways) -------------------------------------------
def _check_tc_bf16(bank_cpu, q_cpu, k, nb="eps", nq="eps"):
    """The bf16 path is exact with respect to the bf16-ROUNDED bank (recall vs the fp32 bank is reported
    separately, tests/test_gpu_parity.py)."""
    fb = FactBank.from_embeddings(bank_cpu, dtype="bf16")
    rounded = bank_cpu.to(torch.bfloat16).float()
    s, i = fb.search(q_cpu.cuda(), k, normalize_bank=nb, normalize_query=nq, flags=L.FORCE_TENSOR)
    kk = min(k, len(bank_cpu))
    truth = ss.truth_scores_fp64(q_cpu, rounded, normalize_bank=bool(nb), normalize_query=bool(nq))
    ss.compare_topk(s.cpu().numpy()[:, :kk], i.cpu().numpy()[:, :kk], truth, k, gap=GAP, rtol=RTOL, atol=ATOL)
    s2, i2 = fb.search(q_cpu.cuda(), k, normalize_bank=nb, normalize_query=nq, flags=L.FORCE_STREAM)
    ss.compare_topk(s2.cpu().numpy()[:, :kk], i2.cpu().numpy()[:, :kk], truth, k, gap=GAP, rtol=RTOL, atol=ATOL)
    fb.close()


@pytest.mark.parametrize("n,d,nq,k", [(128, 64, 16, 4), (1, 8, 3, 2), (129, 72, 9, 8), (4097, 768, 33, 32),
                                       (3000, 4096, 40, 16), (5000, 384, 100, 16), (2000, 200, 17, 128)])
def test_tc_bf16_bank(n, d, nq, k):
    _check_tc_bf16(_randn(n, d, n + d, unit=False), _randn(nq, d, nq + k, unit=False), k)


def test_tc_bf16_hidden_states_and_variants():
    bank, q = _hidden_like(8174, 4096, 2234), _hidden_like(64, 4096, 6678)
    _check_tc_bf16(bank, q, 16)
    _check_tc_bf16(bank[:1000, :512].contiguous(), q[:, :512].contiguous(), 16, nb=None, nq=None)


def test_tc_bf16_is_the_default_for_batches():
    bank, q = _randn(50_000, 768, 5), _randn(64, 768, 6)
    fb = FactBank.from_embeddings(bank, dtype="bf16")
    a = fb.search(q.cuda(), 32)
    b = fb.search(q.cuda(), 32, flags=L.FORCE_TENSOR)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    fb.close()
