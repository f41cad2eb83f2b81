#!/usr/bin/env python
"""bench.py -- IKE top-k retrieval throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (named in config.workload): the north-star configuration -- a 10M x 768 fp32 fact bank,
k = 32, Q = 64 queries per search call (the top of config C3's batch range; tcgen05 kernel) --
row-sharded over the N GPUs (STRONG scaling: the bank is fixed, each GPU scans 10M/N rows).  A
"step" is one search call: fused normalise + contraction + top-k over the whole bank, (N>1: + one
all-gather + merge).  "regimes" adds a short sweep over Q = 1 ... 4096 on the same bank so that the
HBM-bound small-batch kernels and the tensor-core large-batch regime (C4) are reported next to the headline.

value  queries/s with queries and bank resident in HBM, CUDA-event timed, max over ranks.
e2e    the same metric through the host-buffer C-ABI call (numpy queries in, numpy results
       out; H2D + D2H + the sync inside the timed region).
roofline  the scan kernel's algorithmic bytes / its CUDA-event duration vs the measured HBM peak.
cpu_baseline / --impl reference: the reference's CPU path (oracle restatement of
       sentence_transformers.util.semantic_search on torch-CPU, all host threads) on a bounded
       row sample of the same bank shape, extrapolated linearly in rows and labelled as such.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "IKE top-k retrieval queries/s"
UNIT = "queries/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--queries", type=int, default=64)
    ap.add_argument("--topk", type=int, default=32)
    ap.add_argument("--bank-dtype", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--cpu-sample-rows", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-regimes", action="store_true", help="skip the short Q sweep reported under \"regimes\"")
    ap.add_argument("--flags", type=int, default=0, help="extra DEVQA_* flag bits (debugging)")
    return ap.parse_args()


def workload_name(a):
    return (f"{a.rows / 1e6:g}M facts x {a.dim}-d {a.bank_dtype} bank row-sharded over {a.gpus} GPU(s), "
            f"batch {a.queries} queries/call, k={a.topk}, fused eps-normalise (north-star / C3-C4 shape)")


def config_dict(a):
    return {"workload": workload_name(a), "bank_rows": a.rows, "dim": a.dim, "queries_per_step": a.queries,
            "top_k": a.topk, "bank_dtype": a.bank_dtype, "sharding": f"rows/{a.gpus}",
            "l2_policy": "inputs larger than L2 (each step streams the whole shard, >= 3.8 GB >> 126 MB)"}


# ---- CPU reference path (oracle) ------------------------------------------------------------------------
def cpu_reference_run(a, steps, warmup, budget_s):
    """Time the oracle's semantic_search on a bounded row sample; returns (queries/s extrapolated to
    the full bank, dict describing the run)."""
    import torch
    from oracle import semantic_search as ss
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    rows = min(a.cpu_sample_rows, a.rows)
    g = torch.Generator().manual_seed(4234)
    bank = torch.nn.functional.normalize(torch.randn(rows, a.dim, generator=g), dim=1)
    q = torch.nn.functional.normalize(torch.randn(a.queries, a.dim, generator=g), dim=1)

    def one():
        # the dataset/vllm.py pattern: bank normalised once up front, query normalised per call
        return ss.semantic_search(ss.normalize_embeddings(q), bank, score_function=ss.dot_score, top_k=a.topk)

    t0 = time.perf_counter(); one(); t_first = time.perf_counter() - t0
    if budget_s and steps * t_first > budget_s and rows > 100_000:            # keep the run bounded
        rows = max(100_000, int(rows * budget_s / (steps * t_first)))
        bank = bank[:rows].contiguous()
    for _ in range(max(0, warmup - 1)):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = (time.perf_counter() - t0) / steps
    scale = a.rows / rows
    qps = a.queries / (dt * scale)
    info = {"value": qps, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": (f"oracle semantic_search (torch-CPU, {cores} threads), {steps} calls of {a.queries} queries "
                       f"x k={a.topk} over a {rows}-row x {a.dim} fp32 sample ({dt * 1e3:.1f} ms/call), "
                       f"extrapolated linearly x{scale:g} to {a.rows} rows; bank pre-normalised once")}
    return qps, dt * scale * 1e3, info


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    qps, ms, info = cpu_reference_run(a, a.steps, max(a.warmup, 1), budget_s=150.0)
    line = {"impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(a),
            "cpu_baseline": info,
            "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---- clocks ----------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU during the timed region (NVML)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.ok = index, [], False, False
        self.sm_max = None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:          # noqa: BLE001
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:       # noqa: BLE001
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), mhz, reasons))
            except Exception:           # noqa: BLE001
                pass
            time.sleep(0.01)

    def summary(self, t0, t1):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "NVML unavailable: " + getattr(self, "err", "")}
        inside = [s for s in self.samples if t0 <= s[0] <= t1] or self.samples[-3:]
        mhz = sorted(s[1] for s in inside)
        bits = 0
        for s in inside:
            bits |= s[2]
        reasons = sorted({name for bit, name in self.REASONS.items() if bits & bit})
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None, "sm_max_mhz": self.sm_max, "reasons": reasons,
                "samples": len(inside)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:                   # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


def measured_tf32_peak():
    """Dense TF32 tensor throughput to hold the large-batch regimes against: half of the measured SUSTAINED bf16
    cuBLAS figure (TF32 runs at half the bf16 rate; sustained because these are long, power-capped steps)."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["bf16_tflops_sustained"]) / 2.0, "MEASURED_PEAKS.json bf16_tflops_sustained / 2"
    except Exception:                   # noqa: BLE001
        return 1125.0 * 0.62, "fallback: nominal 1125 TF32 TFLOP/s x 0.62 (sustained / nominal bf16 in B200_PROFILING.md)"


def ncu_traffic(a, rows_local, q_per_launch):
    """dram bytes per scan launch from a committed ncu --set full capture of the same shape, else None."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            for rec in json.load(f):
                if (rec["rows"], rec["dim"], rec["queries"], rec["top_k"], rec["bank_dtype"]) == \
                        (rows_local, a.dim, int(q_per_launch), a.topk, a.bank_dtype):
                    return rec["dram_bytes_per_launch"]
    except Exception:                   # noqa: BLE001
        pass
    return None


def kernel_label(nq, bank_dtype, topk):
    """Which scan kernel api.cu selects for this batch (csrc/api.cu run_search)."""
    if nq <= 2:
        return "scan_stream_kernel (CUDA cores, 128-bit streaming loads)"
    if topk <= 64:
        return "scan_tc_kernel<%s filter> (tcgen05/TMEM/TMA) + refine_kernel (exact fp32 re-scoring)" % (
            "TF32x1" if bank_dtype == "fp32" else "bf16x1")
    return "scan_tc_kernel<%s> (tcgen05/TMEM/TMA)" % ("3xTF32" if bank_dtype == "fp32" else "bf16x3")


def filter_fallbacks(L, C, device):
    n = C.c_int64()
    L.check(L.lib().devqa_filter_fallbacks(device, C.byref(n)))
    return int(n.value)


# ---- our arm ------------------------------------------------------------------------------------------------
def run_ours(a):
    import ctypes as C
    import numpy as np
    import torch
    import torch.distributed as dist
    import devqa_b200
    from devqa_b200 import FactBank
    from devqa_b200.dist import Communicator, ShardedFactBank, shard_bounds

    L = devqa_b200._lib
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run for N>1")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        comm = Communicator(device=local_rank)
        bank = ShardedFactBank.synthetic(a.rows, a.dim, comm, seed=4234, dtype=a.bank_dtype)
        local_bank = bank.local
    else:
        bank = local_bank = FactBank(a.dim, device=local_rank, dtype=a.bank_dtype, capacity=a.rows)
        for r0 in range(0, a.rows, 1 << 20):
            bank.append_synthetic(min(1 << 20, a.rows - r0), 4234, first_global_row=r0)
    rows_local = len(local_bank)
    g = torch.Generator().manual_seed(8678)
    q_host = torch.nn.functional.normalize(torch.randn(a.queries, a.dim, generator=g), dim=1)
    q_dev = q_host.to(dev)
    q_np = q_host.numpy()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput ----------------------------------------------------------
    for _ in range(max(a.warmup, 3)):
        out = bank.search(q_dev, a.topk, flags=a.flags)
    barrier()
    filter_fallbacks(L, C, local_rank)                     # reset the counter
    sampler = ClockSampler(local_rank)
    sampler.start()
    L.check(L.lib().devqa_profile_enable(local_rank, 1))
    launches0 = L.lib().devqa_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.perf_counter()
    ev0.record()
    for _ in range(a.steps):
        out = bank.search(q_dev, a.topk, flags=a.flags)
    ev1.record()
    barrier()
    t_wall1 = time.perf_counter()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    launches = L.lib().devqa_launch_count() - launches0
    if world > 1:
        launches += a.steps            # one ncclAllGather kernel per step (NCCL's, counted separately below)
    scan_ms, scan_launches, searches = C.c_double(), C.c_int64(), C.c_int64()
    L.check(L.lib().devqa_profile_collect(local_rank, C.byref(scan_ms), C.byref(scan_launches), C.byref(searches)))
    L.check(L.lib().devqa_profile_enable(local_rank, 0))
    sampler.stop_flag = True
    sampler.join(timeout=2)
    clocks = sampler.summary(t_wall0, t_wall1)
    ms_per_step = ms_total / a.steps
    value = a.queries / (ms_per_step * 1e-3)
    fallbacks = filter_fallbacks(L, C, local_rank)         # queries the exact 3xTF32 fallback had to redo (timed region)

    # ---- end to end through the host-buffer entry point ---------------------------------------
    for _ in range(3):
        bank.search_host(q_np, a.topk)
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        hs, hi = bank.search_host(q_np, a.topk)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = a.queries * a.steps / e2e_s
    same = bool(np.array_equal(hi, out[1].cpu().numpy()) and np.array_equal(hs, out[0].cpu().numpy()))

    # ---- roofline of the dominant kernel (scan) ------------------------------------------------
    esz = 4 if a.bank_dtype == "fp32" else 2
    q_per_launch = a.queries / max(1, scan_launches.value / max(1, searches.value))
    alg_bytes = rows_local * a.dim * esz + q_per_launch * a.dim * 4 + q_per_launch * a.topk * 8
    scan_ms_per_launch = scan_ms.value / max(1, scan_launches.value)
    achieved = alg_bytes / (scan_ms_per_launch * 1e-3) / 1e9 if scan_ms_per_launch > 0 else 0.0
    peak, peak_src = measured_peak()
    kernel_name = kernel_label(a.queries, a.bank_dtype, a.topk)
    roofline = {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": ncu_traffic(a, rows_local, q_per_launch), "peak_source": peak_src,
                "frac_of_8TBs_nominal": achieved / 8000.0, "alg_bytes_per_launch": alg_bytes,
                "launch_ms": scan_ms_per_launch, "launches_timed": scan_launches.value,
                "scan_share_of_step": scan_ms.value / max(1e-9, ev0.elapsed_time(ev1))}

    # ---- regimes: the same bank at other batch sizes (device-resident, 10 steps each) ----------
    regimes = []
    if not a.no_regimes:
        for nq in (1, 4, 16, 64, 128, 1024, 4096):
            gq = torch.Generator().manual_seed(9000 + nq)
            qd = torch.nn.functional.normalize(torch.randn(nq, a.dim, generator=gq), dim=1).to(dev)
            reps = 10 if nq <= 128 else (4 if nq <= 1024 else 2)
            for _ in range(3 if nq <= 128 else 1):
                bank.search(qd, a.topk)
            barrier()
            filter_fallbacks(L, C, local_rank)
            L.check(L.lib().devqa_profile_enable(local_rank, 1))
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record()
            for _ in range(reps):
                bank.search(qd, a.topk)
            r1.record()
            barrier()
            ms = max_over_ranks(r0.elapsed_time(r1)) / reps
            sm, sl, sr = C.c_double(), C.c_int64(), C.c_int64()
            L.check(L.lib().devqa_profile_collect(local_rank, C.byref(sm), C.byref(sl), C.byref(sr)))
            L.check(L.lib().devqa_profile_enable(local_rank, 0))
            scan_ms_call = sm.value / max(1, sr.value)
            gbs = (rows_local * a.dim * esz * (sl.value / max(1, sr.value))) / (scan_ms_call * 1e-3) / 1e9
            passes = sl.value / max(1, sr.value)
            # useful flops of the contraction (2*Q*N*D, split / filter overhead not counted), whole job
            tflops = 2.0 * nq * a.rows * a.dim / (ms * 1e-3) / 1e12
            tf32_peak, tf32_src = measured_tf32_peak()
            regimes.append({"queries_per_call": nq, "ms_per_call": ms, "queries_per_s": nq / (ms * 1e-3),
                            "scan_gbs_per_gpu": gbs, "frac_of_measured_hbm": gbs / peak,
                            "useful_tflops": tflops,
                            "frac_of_tf32_tensor_peak": tflops / (tf32_peak * world), "tf32_peak_source": tf32_src,
                            "kernel": ("stream" if nq <= 2 else ("tcgen05 filter + exact re-scoring" if a.topk <= 64 else "tcgen05")),
                            "bank_passes_per_call": passes,
                            "exact_fallback_queries": filter_fallbacks(L, C, local_rank)})

    if rank == 0:
        cpu = None
        if world == 1 and not a.no_cpu_baseline:
            _, _, cpu = cpu_reference_run(a, steps=40, warmup=2, budget_s=25.0)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config_dict(a),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": a.queries * a.dim * 4,
                        "d2h_bytes_per_step": a.queries * a.topk * 12, "matches_device_path": same},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
                "exact_fallback_queries": fallbacks,
                "aggregate_hbm_gbs": (a.rows * a.dim * esz) / (ms_per_step * 1e-3) / 1e9,
                "aggregate_hbm_frac_of_measured": (a.rows * a.dim * esz) / (ms_per_step * 1e-3) / 1e9 / (peak * world)}
        if regimes:
            line["regimes"] = regimes
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
