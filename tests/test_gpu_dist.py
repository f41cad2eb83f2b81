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
        torch.cuda.synchronize()
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
