"""One C2-shaped search a few times (for an ncu launch list)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, devqa_b200
from devqa_b200 import FactBank
g = torch.Generator().manual_seed(2234)
x = torch.randn(8174, 4096, generator=g) * torch.exp(0.5 * torch.randn(8174, 1, generator=g))
x[:, [1, 4096 // 3, 2048, 4094]] *= 50.0
q = torch.randn(256, 4096, generator=g) * torch.exp(0.5 * torch.randn(256, 1, generator=g))
q[:, [1, 4096 // 3, 2048, 4094]] *= 50.0
fb = FactBank.from_embeddings(x)
qd = q.cuda()
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
for _ in range(4):
    fb.search(qd, 16, flags=flags)
torch.cuda.synchronize()
