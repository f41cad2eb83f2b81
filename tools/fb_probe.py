import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C, torch, devqa_b200
from devqa_b200 import FactBank
L = devqa_b200._lib
rows = 2_000_000
cnt = C.c_int64()
for d in (384, 768):
    fb = FactBank(d, capacity=rows)
    for r0 in range(0, rows, 1 << 20):
        fb.append_synthetic(min(1 << 20, rows - r0), seed=100 + d, first_global_row=r0)
    for nq in (5, 64, 300):
        for k in (1, 5, 16, 32, 33, 64):
            for nb, nqm in (("eps", "eps"), ("noeps", "noeps"), (None, None), ("eps", None)):
                q = torch.randn(nq, d, generator=torch.Generator().manual_seed(1)).cuda()
                L.check(L.lib().devqa_filter_fallbacks(0, C.byref(cnt)))
                fb.search(q, k, normalize_bank=nb, normalize_query=nqm)
                L.check(L.lib().devqa_filter_fallbacks(0, C.byref(cnt)))
                if cnt.value:
                    print(f"d={d} nq={nq} k={k} modes={nb},{nqm}: {cnt.value} fallbacks", flush=True)
    fb.close()
print("done")
