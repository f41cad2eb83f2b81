#!/usr/bin/env python
"""Turn gpurun_out/ ncu artefacts into the small text summaries committed under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_r01.csv  > profiles/r01_launches.txt
    python profiles/summarize.py kernel   gpurun_out/scan_r01.ncu-rep  > profiles/r01_scan_full.txt
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sector_hit_rate.pct", "smsp__cycles_active.avg",
]


def launches(path):
    rows = list(csv.DictReader(l for l in open(path) if not l.startswith("==")))
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        name = r["Kernel Name"]
        agg[name][0] += 1
        agg[name][1] += float(r["Metric Value"].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {len(rows)} launches, gpu__time_duration.sum (ns; cold-cache, serialised: compare SHARES)")
    for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n:5d} launches {ns / 1e3:12.1f} us total {ns / n / 1e3:10.1f} us avg {100 * ns / tot:6.2f}%  {name[:150]}")


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print(f"# {path}: {len(data)} captured launches of {data[0][hdr.index('Kernel Name')][:120]}")
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w:75s} {units[i]:>14s}  " + "  ".join(d[i] for d in data))
    print("# warp stall reasons (warps stalled per issue-active cycle)")
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            print(f"{h:90s} " + "  ".join(d[i] for d in data))


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
