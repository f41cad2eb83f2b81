#!/usr/bin/env python
"""configs jsonl (tools/bench_configs.py) -> the markdown table committed under profiles/.

    python tools/summarize_configs.py gpurun_out/configs_r01c.jsonl > profiles/r01c_configs_1gpu.md
"""
import json
import sys

rows = [json.loads(l) for l in open(sys.argv[1])]
out = ["# BASELINE configs measured on 1x B200 (tools/bench_configs.py --cpu; raw lines: the .jsonl next to this file)\n",
       "GPU = CUDA-event time of the search call (device-resident queries), median; CPU = oracle (torch-CPU restatement of\n"
       "semantic_search, all host threads of the GPU box), best-of-N, bank pre-normalised once / re-normalised per call.\n",
       "| config | shape | cache | GPU ms | GPU q/s | note |\n|---|---|---|---|---|---|"]
for d in rows:
    c = d.get("config")
    if c in ("C1", "C2", "C3") and "ms_median" in d:
        note = d.get("kernel", "") + (" frac of measured HBM %.3f" % d["frac_of_measured"] if "frac_of_measured" in d else "")
        out.append(f"| {c} | {d['n']}x{d['d']}, Q={d['q']}, k={d['k']} | {d.get('cache', 'warm')} | {d['ms_median']:.4f} | {d['qps']:.0f} | {note} |")
out.append("\nEnd to end through the host-buffer C-ABI call (numpy in/out, H2D + D2H + sync inside) and the CPU oracle:\n")
for d in rows:
    if "path" in d:
        extra = f" (re-normalising the bank per call: {d['ms_renorm_per_call']:.3f} ms)" if "ms_renorm_per_call" in d else ""
        out.append(f"* {d['config']} {d['n']}x{d['d']} Q={d['q']} k={d['k']}: {d['path']}: {d['ms']:.4f} ms/call, {d['qps']:.0f} q/s{extra}")
out.append("\nC5 (in-place append sweep, 768-d, k=32):\n\n| N | bank | Q | ms | q/s |\n|---|---|---|---|---|")
for d in rows:
    if d.get("config") == "C5" and "ms_median" in d:
        out.append(f"| {d['n']} | {d['bank']} | {d['q']} | {d['ms_median']:.4f} | {d['qps']:.0f} |")
out.append("\n| C5 N | bf16 recall@32 vs fp32 ids | cumulative append seconds fp32 / bf16 (incl. on-device row synthesis) |\n|---|---|---|")
for d in rows:
    if d.get("config") == "C5" and "bf16_recall_at_32_vs_fp32" in d:
        a = d["append_seconds_cumulative"]
        out.append(f"| {d['n']} | {d['bf16_recall_at_32_vs_fp32']:.4f} | {a['fp32']:.3f} / {a['bf16']:.3f} |")
print("\n".join(out))
