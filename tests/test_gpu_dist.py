"""Two-rank NCCL run of the sharded search (one process per GPU).  Needs >= 2 GPUs."""
import os
import socket

import numpy as np
import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import devqa_b200
        from devqa_b200 import FactBank
        from devqa_b200.dist import Communicator, ShardedFactBank
        comm = Communicator(device=rank)
        n, d, k = 200_003, 768, 32
        sb = ShardedFactBank.synthetic(n, d, comm, seed=3234)
        g = torch.Generator().manual_seed(11)
        q = torch.nn.functional.normalize(torch.randn(9, d, generator=g), dim=1)
        s, i = sb.search(q.cuda(), k)
        hs, hi = sb.search_host(q.numpy(), k)
        assert np.array_equal(s.cpu().numpy(), hs) and np.array_equal(i.cpu().numpy(), hi)
        # single-GPU reference of the same global bank
        full = FactBank(d, device=rank, capacity=n)
        full.append_synthetic(n, seed=3234)
        rs, ri = full.search(q.cuda(), k)
        assert torch.equal(ri, i) and torch.equal(rs, s)
        # chunked, appendable layout
        cb = ShardedFactBank(d, comm, chunk_rows=1024)
        rows = full.rows(0, 10_000)
        for lo, hi_ in [(0, 1), (1, 2), (2, 1500), (1500, 10_000)]:
            cb.append(rows[lo:hi_])
        cs, ci = cb.search(q.cuda(), k)
        small = FactBank.from_embeddings(rows, device=rank)
        ss_, si_ = small.search(q.cuda(), k)
        assert torch.equal(ci, si_) and torch.equal(cs, ss_)
        # ---- sharded bank with a bf16 shadow: same results, bit for bit (the shards stream 2 bytes per element)
        sbs = ShardedFactBank.synthetic(n, d, comm, seed=3234, shadow=True)
        assert sbs.local.has_shadow
        q40 = torch.nn.functional.normalize(torch.randn(40, d, generator=g), dim=1).cuda()
        ss_, si_ = sbs.search(q40, k)
        rs2, ri2 = full.search(q40, k)
        assert torch.equal(si_, ri2) and torch.equal(ss_, rs2)
        # ---- the exchange: peer-to-peer mailboxes (default) vs ncclAllGather, bit for bit; batches beyond the mailbox
        # (Q > 512) take the NCCL path by themselves
        L = devqa_b200._lib
        assert comm.p2p == bool(L.lib().devqa_dist_p2p_enabled(comm._h))
        report = {"p2p": comm.p2p}
        big_q = torch.nn.functional.normalize(torch.randn(600, d, generator=g), dim=1).cuda()
        for nq, kk in ((1, 5), (9, 32), (64, 32), (300, 100), (512, 128)):
            qq = big_q[:nq]
            a = sb.search(qq, kk)
            if comm.p2p:
                L.check(L.lib().devqa_dist_p2p_set(comm._h, 0))
                b = sb.search(qq, kk)
                L.check(L.lib().devqa_dist_p2p_set(comm._h, 1))
                assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]), (nq, kk)
            r = full.search(qq, kk)
            assert torch.equal(a[1], r[1]) and torch.equal(a[0], r[0]), (nq, kk)
        a = sb.search(big_q, 8)
        r = full.search(big_q, 8)
        assert torch.equal(a[1], r[1]) and torch.equal(a[0], r[0])
        # ---- the whole distributed step (scan -> re-scoring -> exchange -> merge) captured in ONE CUDA graph and
        # replayed with new queries: no host involvement per step (missing #5)
        qbuf = big_q[:64].clone()
        out_s = torch.empty((64, k), dtype=torch.float32, device="cuda")
        out_i = torch.empty((64, k), dtype=torch.int64, device="cuda")
        from devqa_b200.bank import norm_flags, _stream_ptr

        def enqueue():
            L.check(L.lib().devqa_bank_search_dist(sb.local._h, comm._h, qbuf.data_ptr(), 64, d, k, norm_flags("eps", "eps"),
                                                   out_s.data_ptr(), out_i.data_ptr(), _stream_ptr(rank)))
        enqueue(); enqueue()                                        # warm-up: sizes every scratch buffer
        torch.cuda.synchronize()
        dist.barrier()
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.graph(graph, stream=side):
            enqueue()
        for rep in range(4):
            qbuf.copy_(big_q[rep * 100: rep * 100 + 64])
            graph.replay()
            torch.cuda.synchronize()
            want = full.search(qbuf, k)
            assert torch.equal(out_i, want[1]) and torch.equal(out_s, want[0]), rep
        # timing of the eager step, both exchanges (reported, not asserted)
        for name, on in (("p2p", 1), ("nccl", 0)):
            if not comm.p2p and on:
                continue
            if comm.p2p:
                L.check(L.lib().devqa_dist_p2p_set(comm._h, on))
            for _ in range(5):
                sb.search(qbuf, k)
            dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(100):
                sb.search(qbuf, k)
            e1.record(); torch.cuda.synchronize()
            report[name + "_us_per_step"] = e0.elapsed_time(e1) * 10.0
        if comm.p2p:
            L.check(L.lib().devqa_dist_p2p_set(comm._h, 1))
        assert comm.p2p_timeouts() == 0
        torch.cuda.synchronize()
        import json
        gp = os.path.join(ROOT, "gpurun_out")
        if rank == 0 and os.path.isdir(gp):
            json.dump(report, open(os.path.join(gp, "dist_exchange_2gpu.json"), "w"))
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_gpu_sharded_search(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(2))


def test_bank_on_device_1_while_device_0_is_current():
    """the per-device scratch (its event in particular) must be created on the bank's device, whichever
    device the calling thread has current -- bank_from_file(device=hparams.device) is called exactly like that."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import subprocess, sys, textwrap
    code = textwrap.dedent(f"""
        import sys; sys.path.insert(0, {ROOT!r})
        import torch, devqa_b200
        from devqa_b200 import FactBank, EditPool, LiveEditPool
        torch.cuda.set_device(0)
        g = torch.Generator().manual_seed(1)
        bank = torch.randn(5000, 128, generator=g)
        q = torch.randn(9, 128, generator=g)
        fb1 = FactBank.from_embeddings(bank, device=1)          # first touch of device 1's scratch, device 0 current
        s1, i1 = fb1.search(q, 8)
        side = torch.cuda.Stream(device=1)
        with torch.cuda.stream(side):                           # stream change -> the scratch's event is waited on
            s1b, i1b = fb1.search(q, 8)
        fb0 = FactBank.from_embeddings(bank, device=0)
        s0, i0 = fb0.search(q, 8)
        torch.cuda.synchronize(0); torch.cuda.synchronize(1)
        assert s1.device.index == 1 and torch.equal(i1.cpu(), i0.cpu()) and torch.equal(s1.cpu(), s0.cpu())
        assert torch.equal(i1b.cpu(), i0.cpu())
        pool = EditPool(128, device=1); pool.append(bank[:100])
        ids, (sims, order) = pool.recipe_retrieve(q, 4)
        lp = LiveEditPool(4, 32, device=1); lp.append(bank[:50].reshape(50, 4, 32))
        lp.retrieve(q[:1].reshape(1, 4, 32), torch.tensor([[0.0]]), 10.0)
        print("ok")
    """)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]
