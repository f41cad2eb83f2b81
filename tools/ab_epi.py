"""A/B inside one process (alternating): scan variants selected by debug bits, same bank and queries.

  no_resident  bit 6   the query operand streamed with the ring stages (default: resident in shared memory when it fits)
  no_evict_first  bit 5   bank tiles loaded without the L2 evict-first hint
  one_group  bit 4   emit mode with one group of epilogue warps (default: two at 128+ queries per pass)
  pair_all   bit 29  CTA pairs also at 64 / 128 queries per pass (default: only at 256)
Also checks that all variants return the same ids and scores (and, on small banks, the same as the stream kernel)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, devqa_b200
from devqa_b200 import FactBank
L = devqa_b200._lib
DIM = int(os.environ.get('AB_DIM', '768'))
ROWS = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
QS = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [64, 128, 256, 1024, 4096]
SHADOWS = [x == "1" for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [True, False]
VARIANTS = {"default": 0, "no_resident": 1 << 6}
if os.environ.get("AB_EVICT"):
    VARIANTS["no_evict_first"] = 1 << 5
if os.environ.get("AB_ALL"):
    VARIANTS.update({"one_group": 1 << 4, "pair_all": 1 << 29})
def fallbacks():
    v = C.c_int64(0)
    L.check(L.lib().devqa_filter_fallbacks(0, C.byref(v)))
    return v.value
for shadow in SHADOWS:
    fb = FactBank(DIM, capacity=ROWS, shadow=shadow)
    for r0 in range(0, ROWS, 1 << 20):
        fb.append_synthetic(min(1 << 20, ROWS - r0), 4234, first_global_row=r0)
    g = torch.Generator().manual_seed(8678)
    for nq in QS:
        q = torch.nn.functional.normalize(torch.randn(nq, DIM, generator=g), dim=1).cuda()
        outs = {k: fb.search(q, 32, flags=f) for k, f in VARIANTS.items()}
        if nq <= 256 and ROWS <= 2_000_000:
            outs["stream"] = fb.search(q, 32, flags=L.FORCE_STREAM)
        a = outs["default"]
        same = all(bool((a[1] == b[1]).all()) and bool((a[0] == b[0]).all()) for b in outs.values())
        res = {k: [] for k in VARIANTS}
        reps = 3 if nq >= 1024 else 5
        for rep in range(reps):
            for k, flags in VARIANTS.items():
                fb.search(q, 32, flags=flags)
                torch.cuda.synchronize()
                n = 3 if nq >= 1024 else 10
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(n):
                    fb.search(q, 32, flags=flags)
                e1.record(); torch.cuda.synchronize()
                res[k].append(e0.elapsed_time(e1) / n)
        med = {k: sorted(v)[reps // 2] for k, v in res.items()}
        print(f"shadow={int(shadow)} rows={ROWS} Q={nq}: " + "  ".join(f"{k} {m:.3f} ms ({nq / m:.1f} k q/s)" for k, m in med.items()) +
              f"  identical={same} fallbacks={fallbacks()}", flush=True)
    del fb
    torch.cuda.empty_cache()
