"""C2 (8174 x 4096, Q = 256, k = 16) timed with CUDA events: one-shot mode vs the persistent list epilogue."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, devqa_b200
from devqa_b200 import FactBank
L = devqa_b200._lib
fb = FactBank(4096, capacity=8174, shadow="--shadow" in sys.argv)
fb.append_synthetic(8174, 2234, kind=1)
g = torch.Generator().manual_seed(6678)
q = (torch.randn(256, 4096, generator=g) * torch.exp(0.5 * torch.randn(256, 1, generator=g))).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(flags, cold, iters=30, nq=256, k=16):
    qq = q[:nq]
    for _ in range(5):
        fb.search(qq, k, flags=flags)
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if cold:
            flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fb.search(qq, k, flags=flags); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[0], ts[len(ts) // 2]


for name, flags in (("one-shot", 0), ("no shadow", L.NO_SHADOW), ("list epilogue", L.DEBUG_NO_ONESHOT)):
    for nq in (256, 64, 16):
        print(f"{name:14s} Q={nq:4d} cold best/median us {timed(flags, True, nq=nq)}  warm {timed(flags, False, nq=nq)}", flush=True)
a = fb.search(q, 16)
b = fb.search(q, 16, flags=L.DEBUG_NO_ONESHOT)
c = fb.search(q, 16, flags=L.NO_SHADOW)
print("identical:", torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]), torch.equal(a[0], c[0]) and torch.equal(a[1], c[1]))
