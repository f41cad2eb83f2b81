"""Fixed overhead of a search: time Q=64, k=32 over a 768-d bank truncated to several sizes (one process)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, devqa_b200
from devqa_b200 import FactBank
L = devqa_b200._lib
shadow = "--no-shadow" not in sys.argv
nq = 64
fb = FactBank(768, capacity=10_000_000, shadow=shadow)
for r0 in range(0, 10_000_000, 1 << 20):
    fb.append_synthetic(min(1 << 20, 10_000_000 - r0), 4234, first_global_row=r0)
g = torch.Generator().manual_seed(8678)
q = torch.nn.functional.normalize(torch.randn(nq, 768, generator=g), dim=1).cuda()
print("shadow", shadow)
for rows in (10_000_000, 5_000_000, 2_500_000, 1_250_000, 625_000, 312_500, 156_250):
    fb.truncate(rows)
    for _ in range(5):
        fb.search(q, 32)
    torch.cuda.synchronize()
    L.check(L.lib().devqa_profile_enable(0, 1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 50
    e0.record()
    for _ in range(reps):
        fb.search(q, 32)
    e1.record(); torch.cuda.synchronize()
    sm, sl, sr = C.c_double(), C.c_int64(), C.c_int64()
    L.check(L.lib().devqa_profile_collect(0, C.byref(sm), C.byref(sl), C.byref(sr)))
    L.check(L.lib().devqa_profile_enable(0, 0))
    ms = e0.elapsed_time(e1) / reps
    scan = sm.value / reps
    gb = rows * 768 * (2 if shadow else 4) / 1e9
    print(f"rows {rows:9d}  step {ms*1e3:8.1f} us  scan {scan*1e3:8.1f} us  non-scan {(ms-scan)*1e3:6.1f} us  scan GB/s {gb/scan*1e3:7.0f}", flush=True)
