"""C1-shaped searches (6345 x 384, Q = 1, k = 16) for an ncu launch list / latency probe."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, devqa_b200
from devqa_b200 import FactBank
g = torch.Generator().manual_seed(1234)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6345
bank = torch.nn.functional.normalize(torch.randn(n, 384, generator=g), dim=1)
q = torch.nn.functional.normalize(torch.randn(64, 384, generator=g), dim=1)
fb = FactBank.from_embeddings(bank)
qd = q.cuda()
for i in range(20):
    fb.search(qd[i:i + 1], 16)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(200):
    fb.search(qd[i % 64:i % 64 + 1], 16)
e1.record(); e1.synchronize()
print("device us/search (back to back)", e0.elapsed_time(e1) * 1e3 / 200)
qn = q.numpy()
t0 = time.perf_counter()
for i in range(500):
    fb.search_host(qn[i % 64], 16)
print("e2e us/search", (time.perf_counter() - t0) / 500 * 1e6)
