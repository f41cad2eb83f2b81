"""A/B inside one process: emit-mode scan with 8 accumulator buffers vs 2 (debug bit 23), alternating."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, devqa_b200
from devqa_b200 import FactBank
L = devqa_b200._lib
fb = FactBank(768, capacity=2_500_000, shadow=True)
for r0 in range(0, 2_500_000, 1 << 20):
    fb.append_synthetic(min(1 << 20, 2_500_000 - r0), 4234, first_global_row=r0)
g = torch.Generator().manual_seed(8678)
q = torch.nn.functional.normalize(torch.randn(64, 768, generator=g), dim=1).cuda()
for rows in (2_500_000, 1_250_000, 625_000):
    fb.truncate(rows)
    res = {0: [], 1 << 23: []}
    for rep in range(6):
        for flags in (0, 1 << 23):
            for _ in range(3):
                fb.search(q, 32, flags=flags)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(40):
                fb.search(q, 32, flags=flags)
            e1.record(); torch.cuda.synchronize()
            res[flags].append(e0.elapsed_time(e1) / 40 * 1e3)
    print(rows, "8 buffers: %.1f us (min %.1f)   2 buffers: %.1f us (min %.1f)" % (
        sorted(res[0])[3], min(res[0]), sorted(res[1 << 23])[3], min(res[1 << 23])), flush=True)
