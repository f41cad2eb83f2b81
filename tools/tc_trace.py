"""Print the per-role clock64 timeline of CTA 0 of the tcgen05 scan (debugging aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, devqa_b200
L = devqa_b200._lib
rows, q, flags = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 0
dtype = sys.argv[4] if len(sys.argv) > 4 else "fp32"
shadow = dtype == "shadow"
fb = devqa_b200.FactBank(768, capacity=rows, dtype="fp32" if shadow else dtype, shadow=shadow)
fb.append_synthetic(rows, 1)
qs = torch.nn.functional.normalize(torch.randn(q, 768), dim=1).cuda()
for _ in range(2):
    fb.search(qs, 32, flags=L.FORCE_TENSOR | flags)
buf = torch.zeros(16 * 256, dtype=torch.int64, device="cuda")
L.check(L.lib().devqa_debug_tc_trace(buf.data_ptr()))
fb.search(qs, 32, flags=L.FORCE_TENSOR | flags)
torch.cuda.synchronize()
L.lib().devqa_debug_tc_trace(None)
t = buf.cpu().view(16, 256)
t0 = int(t[0, 0])
names = ["tma_issue", "got_empty", "split_done", "mma_got_tma", "mma_got_peer", "mma_commit", "epi_got", "epi_done"]
n = 60
for r in range(6):
    print(names[r].ljust(14), " ".join(f"{int(x) - t0:7d}" if x else "      -" for x in t[r, :n]))
for r in (6, 7):
    print(names[r].ljust(14), " ".join(f"{int(x) - t0:7d}" if x else "      -" for x in t[r, :24]))
print("slow_calls_w0".ljust(14), " ".join(f"{int(x):7d}" for x in t[8, :24]))
for r, nm in ((9, "passA_done"), (10, "published"), (11, "thr_arrived"), (12, "svc_all_pub"), (13, "svc_selected"), (14, "svc_all_sel")):
    print(nm.ljust(14), " ".join(f"{int(x) - t0:7d}" if x else "      -" for x in t[r, :2]))
# summary over all recorded tiles
import numpy as np
got, done, slow = t[6].numpy().astype(np.int64), t[7].numpy().astype(np.int64), t[8].numpy()
m = (got > 0) & (done > 0)
n_rec = int(m.sum())
if n_rec > 8:
    dur = (done - got)[m]
    per = np.diff(got[m])
    for lo, hi in ((0, 8), (8, 64), (64, 128), (128, 256)):
        if lo < n_rec:
            h = min(hi, n_rec)
            print(f"tiles {lo:3d}-{h:3d}: epilogue cycles/tile {dur[lo:h].mean():8.0f}  tile period {per[lo:h - 1].mean() if h - 1 > lo else 0:8.0f}  slow calls (warp 0)/tile {slow[lo:h].mean():5.2f}")
    print(f"recorded tiles {n_rec}: first accumulator ready at {int(got[m][0]) - t0}, last epilogue done at {int(done[m][-1]) - t0} cycles after the producer's first issue")
    print("first tiles: epilogue cycles", " ".join(str(int(x)) for x in dur[:12]), "| slow calls", " ".join(str(int(x)) for x in slow[:12]))
