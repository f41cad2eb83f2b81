"""Per-call wall-clock of the host-buffer entry point next to the device-timed search (10M x 768, fp16 shadow, Q = 64)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, devqa_b200
from devqa_b200 import FactBank
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
fb = FactBank(768, capacity=rows, shadow=True)
for r0 in range(0, rows, 1 << 20):
    fb.append_synthetic(min(1 << 20, rows - r0), 4234, first_global_row=r0)
q = torch.nn.functional.normalize(torch.randn(64, 768, generator=torch.Generator().manual_seed(1)), dim=1)
q_np, q_dev = q.numpy().copy(), q.cuda()
for _ in range(5):
    fb.search_host(q_np, 32); fb.search(q_dev, 32)
torch.cuda.synchronize()
host = []
for _ in range(40):
    t0 = time.perf_counter(); fb.search_host(q_np, 32); host.append((time.perf_counter() - t0) * 1e3)
dev = []
for _ in range(40):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fb.search(q_dev, 32); e1.record(); torch.cuda.synchronize(); dev.append(e0.elapsed_time(e1))
sync = []
for _ in range(40):
    t0 = time.perf_counter(); fb.search(q_dev, 32); torch.cuda.synchronize(); sync.append((time.perf_counter() - t0) * 1e3)
f = lambda v: f"min {min(v):.3f} median {sorted(v)[len(v)//2]:.3f} max {max(v):.3f}"
print("search_host wall ms:", f(host)); print("device events ms:  ", f(dev)); print("search + sync wall:", f(sync))
print("host calls:", " ".join(f"{x:.2f}" for x in host[:20]))
