"""One search at a given shape, for `ncu` launch lists: python tools/one_call.py ROWS Q [shadow=1] [flags=0] [reps=2]."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, devqa_b200
from devqa_b200 import FactBank
rows, nq = int(sys.argv[1]), int(sys.argv[2])
shadow = (sys.argv[3] if len(sys.argv) > 3 else "1") == "1"
flags = int(sys.argv[4]) if len(sys.argv) > 4 else 0
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
fb = FactBank(768, capacity=rows, shadow=shadow)
for r0 in range(0, rows, 1 << 20):
    fb.append_synthetic(min(1 << 20, rows - r0), 4234, first_global_row=r0)
q = torch.nn.functional.normalize(torch.randn(nq, 768, generator=torch.Generator().manual_seed(8678)), dim=1).cuda()
for _ in range(reps):
    fb.search(q, 32, flags=flags)
torch.cuda.synchronize()
