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
# (round 2: each width once with and once without the bf16 shadow; sizes that take the emit, one-shot and list paths)
for d, shadow, n in ((384, True, rows), (768, False, rows), (768, True, rows // 4), (256, False, 15_000), (1024, True, 100_000)):
    fb = FactBank(d, capacity=n, shadow=shadow)
    for r0 in range(0, n, 1 << 20):
        fb.append_synthetic(min(1 << 20, n - r0), seed=100 + d, first_global_row=r0, kind=1 if d == 1024 else 0)
    banks[(d, shadow, n)] = fb
cnt = C.c_int64()
L.check(L.lib().devqa_filter_fallbacks(0, C.byref(cnt)))
t0 = time.time()
fallbacks = 0
per_bank = {}
for it in range(iters):
    key = rnd.choice(list(banks))
    d, rows_now = key[0], key[2]
    fb = banks[key]
    nq = rnd.choice([3, 4, 5, 16, 17, 63, 64, 65, 100, 128, 129, 200, 256, 257, 300, 512, 600, 1025])
    k = rnd.choice([1, 5, 16, 32, 33, 64, 100, 128])
    nb, nqm = rnd.choice([("eps", "eps"), ("noeps", "noeps"), (None, None), ("eps", None)])
    g = torch.Generator().manual_seed(it)
    q = torch.randn(nq, d, generator=g).cuda()
    if rnd.random() < 0.3:                                    # some queries that ARE bank rows
        ids = torch.randint(0, rows_now, (min(4, nq),), generator=g)
        for j, r in enumerate(ids.tolist()):
            q[j] = fb.rows(r, r + 1)[0]
    s, i = fb.search(q, k, normalize_bank=nb, normalize_query=nqm)
    sel = torch.tensor(sorted(rnd.sample(range(nq), min(2, nq))))
    s2, i2 = fb.search(q[sel], k, normalize_bank=nb, normalize_query=nqm, flags=L.FORCE_STREAM)
    L.check(L.lib().devqa_filter_fallbacks(0, C.byref(cnt)))
    fallbacks += cnt.value
    per_bank[key] = per_bank.get(key, [0, 0])
    per_bank[key][0] += cnt.value
    per_bank[key][1] += nq + len(sel)
    if not (torch.equal(i[sel], i2) and torch.equal(s[sel], s2)):
        print("MISMATCH", it, key, nq, k, nb, nqm, flush=True)
        sys.exit(1)
    assert bool((s[:, 1:] <= s[:, :-1]).all() | torch.isnan(s).any())
    if it % 50 == 49:
        print(f"{it + 1} searches ok, {fallbacks} fallback queries, {time.time() - t0:.1f} s", flush=True)
L.check(L.lib().devqa_filter_bound_violations(0, C.byref(cnt)))
print("soak ok:", iters, "searches,", fallbacks, "fallback queries,", cnt.value, "bound violations")
assert cnt.value == 0
for key, (fbq, total) in per_bank.items():
    print(f"  bank dim {key[0]} shadow {key[1]} rows {key[2]}: {fbq} fallback queries of {total}")
