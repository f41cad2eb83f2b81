"""Soak test of the tcgen05 paths: many searches with random batch sizes / k / modes on a mid-size and a large bank,
every result compared bit for bit with the CUDA-core stream kernel on a sample of its queries.

    python tools/soak.py [iterations] [rows]
"""
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C

import torch

import devqa_b200
from devqa_b200 import FactBank

L = devqa_b200._lib
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
rnd = random.Random(7)
banks = {}
for d in (384, 768):
    fb = FactBank(d, capacity=rows)
    for r0 in range(0, rows, 1 << 20):
        fb.append_synthetic(min(1 << 20, rows - r0), seed=100 + d, first_global_row=r0)
    banks[d] = fb
cnt = C.c_int64()
L.check(L.lib().devqa_filter_fallbacks(0, C.byref(cnt)))
t0 = time.time()
fallbacks = 0
for it in range(iters):
    d = rnd.choice([384, 768])
    fb = banks[d]
    nq = rnd.choice([3, 4, 5, 16, 17, 63, 64, 65, 100, 128, 129, 200, 256, 257, 300, 512, 600, 1025])
    k = rnd.choice([1, 5, 16, 32, 33, 64])
    nb, nqm = rnd.choice([("eps", "eps"), ("noeps", "noeps"), (None, None), ("eps", None)])
    g = torch.Generator().manual_seed(it)
    q = torch.randn(nq, d, generator=g).cuda()
    if rnd.random() < 0.3:                                    # some queries that ARE bank rows
        ids = torch.randint(0, rows, (min(4, nq),), generator=g)
        for j, r in enumerate(ids.tolist()):
            q[j] = fb.rows(r, r + 1)[0]
    s, i = fb.search(q, k, normalize_bank=nb, normalize_query=nqm)
    sel = torch.tensor(sorted(rnd.sample(range(nq), min(2, nq))))
    s2, i2 = fb.search(q[sel], k, normalize_bank=nb, normalize_query=nqm, flags=L.FORCE_STREAM)
    L.check(L.lib().devqa_filter_fallbacks(0, C.byref(cnt)))
    fallbacks += cnt.value
    if not (torch.equal(i[sel], i2) and torch.equal(s[sel], s2)):
        print("MISMATCH", it, d, nq, k, nb, nqm, flush=True)
        sys.exit(1)
    assert bool((s[:, 1:] <= s[:, :-1]).all() | torch.isnan(s).any())
    if it % 50 == 49:
        print(f"{it + 1} searches ok, {fallbacks} fallback queries, {time.time() - t0:.1f} s", flush=True)
print("soak ok:", iters, "searches,", fallbacks, "fallback queries")
