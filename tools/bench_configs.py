#!/usr/bin/env python
"""Measure every BASELINE.json configuration on one B200 (C1, C2, C3, C5; C4 is bench.py --gpus 8).

    python tools/bench_configs.py [--cpu] > gpurun_out/configs.jsonl

One JSON line per measurement.  GPU times are CUDA-event times of the search call (device-resident
queries); "cold" flushes L2 (256 MB write) between iterations for banks smaller than L2, "warm" does
not.  "e2e" goes through the host-buffer C-ABI entry point.  --cpu adds the oracle (reference CPU
path, all host threads) beside each row, on the same shapes (bounded samples where noted).
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

import devqa_b200
from devqa_b200 import FactBank

L = devqa_b200._lib
PEAK = 6456.8


def unit(n, d, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=1)


def hidden_like(n, d, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, d, generator=g) * torch.exp(0.5 * torch.randn(n, 1, generator=g))
    x[:, [1, d // 3, d // 2, d - 2]] *= 50.0
    return x


_flush = None


def time_search(fn, iters, cold):
    global _flush
    if cold and _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(iters):
        if cold:
            _flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        times.append(e0.elapsed_time(e1))
    times.sort()
    return times[0], times[len(times) // 2]


def emit(**kw):
    print(json.dumps(kw), flush=True)


def cpu_time(q, bank, k, nb, reps=5):
    from oracle import semantic_search as ss
    torch.set_num_threads(os.cpu_count())
    norm = {"eps": ss.normalize_embeddings, "noeps": ss.normalize_noeps, None: (lambda t: t)}
    pre = norm[nb](bank)                                     # dataset/vllm.py pattern: bank normalised once
    best = best_renorm = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        ss.semantic_search(norm[nb](q), pre, score_function=ss.dot_score, top_k=k)
        best = min(best, time.perf_counter() - t0)
        t0 = time.perf_counter()                             # select_icl_samples pattern: renormalise per call
        ss.semantic_search(norm[nb](q), norm[nb](bank), score_function=ss.dot_score, top_k=k)
        best_renorm = min(best_renorm, time.perf_counter() - t0)
    return best * 1e3, best_renorm * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    want = set(a.only.split(",")) if a.only else None
    cores = os.cpu_count()

    # ---- C1: EVQA-shaped bank, one query at a time --------------------------------------------
    if not want or "C1" in want:
        for n in (6345, 19035):
            bank = unit(n, 384, 1234)
            q = unit(64, 384, 5678)
            fb = FactBank.from_embeddings(bank)
            qd = q.cuda()
            for k in (16, 5):
                for cold in (False, True):
                    best, med = time_search(lambda: fb.search(qd[:1], k), 200, cold)
                    emit(config="C1", n=n, d=384, q=1, k=k, cache="cold" if cold else "warm", ms_best=best, ms_median=med,
                         qps=1e3 / med, alg_mb=n * 384 * 4 / 1e6)
            t0 = time.perf_counter()
            for i in range(500):
                fb.search_host(q[i % 64].numpy(), 16)
            dt = (time.perf_counter() - t0) / 500
            emit(config="C1", n=n, d=384, q=1, k=16, path="e2e host buffers", ms=dt * 1e3, qps=1 / dt)
            if a.cpu:
                ms, ms_r = cpu_time(q[:1], bank, 16, "eps", reps=20)
                emit(config="C1", n=n, d=384, q=1, k=16, path=f"cpu oracle {cores} threads", ms=ms, qps=1e3 / ms,
                     ms_renorm_per_call=ms_r)
            fb.close()

    # ---- C2: VLKEB-shaped, D = 4096 hidden-state-like, Q = 256 ----------------------------------
    if not want or "C2" in want:
        bank = hidden_like(8174, 4096, 2234)
        q = hidden_like(256, 4096, 6678)
        fb = FactBank.from_embeddings(bank)
        qd = q.cuda()
        for cold in (False, True):
            best, med = time_search(lambda: fb.search(qd, 16), 30, cold)
            flops = 2 * 256 * 8174 * 4096
            emit(config="C2", n=8174, d=4096, q=256, k=16, cache="cold" if cold else "warm", ms_best=best, ms_median=med,
                 qps=256e3 / med, tflops=flops / (med * 1e-3) / 1e12, kernel="tcgen05 (K-set accumulation)")
        best, med = time_search(lambda: fb.search(qd, 16, flags=L.FORCE_STREAM), 10, False)
        emit(config="C2", n=8174, d=4096, q=256, k=16, cache="warm", ms_median=med, qps=256e3 / med, kernel="stream (forced)")
        t0 = time.perf_counter()
        for _ in range(20):
            fb.search_host(q.numpy(), 16)
        dt = (time.perf_counter() - t0) / 20
        emit(config="C2", n=8174, d=4096, q=256, k=16, path="e2e host buffers", ms=dt * 1e3, qps=256 / dt)
        if a.cpu:
            ms, ms_r = cpu_time(q, bank, 16, "eps", reps=3)
            emit(config="C2", n=8174, d=4096, q=256, k=16, path=f"cpu oracle {cores} threads", ms=ms, qps=256e3 / ms,
                 ms_renorm_per_call=ms_r)
        fb.close()

    # ---- C3: 1M x 768, Q = 1..64, k = 32 ---------------------------------------------------------
    if not want or "C3" in want:
        fb = FactBank(768, capacity=1_000_000)
        fb.append_synthetic(1_000_000, 3234)
        for nq in (1, 2, 4, 8, 16, 32, 64):
            qd = unit(nq, 768, 7678 + nq).cuda()
            best, med = time_search(lambda: fb.search(qd, 32), 50, False)
            emit(config="C3", n=1_000_000, d=768, q=nq, k=32, ms_best=best, ms_median=med, qps=nq * 1e3 / med,
                 gbs=3.072 / (med * 1e-3), frac_of_measured=3.072 / (med * 1e-3) / PEAK,
                 kernel="tcgen05" if nq > 2 else "stream")
        if a.cpu:
            bank = fb.rows().cpu()
            for nq in (1, 8, 64):
                ms, ms_r = cpu_time(unit(nq, 768, 7678 + nq), bank, 32, "eps", reps=3)
                emit(config="C3", n=1_000_000, d=768, q=nq, k=32, path=f"cpu oracle {cores} threads", ms=ms,
                     qps=nq * 1e3 / ms, ms_renorm_per_call=ms_r)
        fb.close()

    # ---- C5: lifelong sweep: in-place append 1k -> 10M, fp32 vs bf16 ------------------------------
    if not want or "C5" in want:
        banks = {"fp32": FactBank(768), "bf16": FactBank(768, dtype="bf16")}
        q1, q64 = unit(1, 768, 9678).cuda(), unit(64, 768, 9679).cuda()
        n = 0
        marks = [1_000, 10_000, 100_000, 1_000_000, 10_000_000]
        append_s = {k: 0.0 for k in banks}
        for mark in marks:
            for name, fb in banks.items():
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                cur = n
                while cur < mark:                             # 1 row at a time to 1k, then 1k batches, then 64k batches
                    step = 1 if cur < 1_000 else (1_000 if cur < 1_000_000 else 65_536)
                    step = min(step, mark - cur)
                    fb.append_synthetic(step, 5234, first_global_row=cur)
                    cur += step
                torch.cuda.synchronize()
                append_s[name] += time.perf_counter() - t0
            rows_added = mark - n
            n = mark
            res = {}
            for name, fb in banks.items():
                for nq, qd in ((1, q1), (64, q64)):
                    best, med = time_search(lambda: fb.search(qd, 32), 30 if n <= 1_000_000 else 10, False)
                    res[(name, nq)] = med
                    emit(config="C5", n=n, d=768, q=nq, k=32, bank=name, ms_median=med, qps=nq * 1e3 / med,
                         capacity=fb.capacity)
            i32 = banks["fp32"].search(q64, 32)[1].cpu().tolist()
            i16 = banks["bf16"].search(q64, 32)[1].cpu().tolist()
            recall = float(np.mean([len(set(x) & set(y)) / min(32, n) for x, y in zip(i32, i16)]))
            emit(config="C5", n=n, bf16_recall_at_32_vs_fp32=recall, rows_appended=rows_added,
                 append_seconds_cumulative=append_s)
        for fb in banks.values():
            fb.close()


if __name__ == "__main__":
    main()
