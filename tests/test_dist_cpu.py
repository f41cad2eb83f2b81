"""World-size-2 gloo test of the N>1 protocol on the CPU: contiguous and chunked sharding, the
packed-key exchange and the order-independent merge.  The device kernels are
replaced by the oracle per shard; what is checked is the host-side logic the GPU path shares:
partition arithmetic, global ids, key format, one all-gather, merged result == unsharded result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import devqa_b200
        from devqa_b200 import dist as dd
        from oracle import semantic_search as ss
        import keyfmt

        g = torch.Generator().manual_seed(99)
        bank = torch.nn.functional.normalize(torch.randn(1003, 32, generator=g), dim=1)
        queries = torch.randn(5, 32, generator=g)
        k = 8
        truth = ss.truth_scores_fp64(queries, bank)

        # ---- contiguous shards -------------------------------------------------------------
        b = dd.shard_bounds(len(bank), world)
        lo, hi = b[rank], b[rank + 1]
        s, i = ss.search_arrays(queries, bank[lo:hi], k, "eps", "eps")
        keys = torch.from_numpy(keyfmt.pack_keys(s, i + lo).astype(np.int64))      # gloo has no uint64
        gathered = [torch.empty_like(keys) for _ in range(world)]
        dist.all_gather(gathered, keys)
        merged = keyfmt.merge_keys(torch.stack(gathered).numpy().view(np.uint64), k)
        ms, mi = keyfmt.unpack_keys(merged)
        ss.compare_topk(ms, mi, truth, k)
        # every rank ends with the identical result
        box = [None] * world
        dist.all_gather_object(box, merged.tobytes())
        assert all(x == box[0] for x in box)

        # ---- chunked (appendable) layout ----------------------------------------------------
        chunk = 64
        local_rows, local_gids, total = [], [], 0
        for batch in [1, 1, 30, 100, 64, 500, 307]:
            rows = bank[total:total + batch]
            for off, cnt in dd.split_append(total, batch, chunk, rank, world):
                local_rows.append(rows[off:off + cnt])
                local_gids += list(range(total + off, total + off + cnt))
            total += batch
        assert total == len(bank)
        mine = torch.cat(local_rows) if local_rows else bank[:0]
        assert [dd.local_to_global(j, chunk, rank, world) for j in range(len(mine))] == local_gids
        s, i = ss.search_arrays(queries, mine, k, "eps", "eps")
        gid = np.asarray(local_gids, dtype=np.int64)[i]
        keys = torch.from_numpy(keyfmt.pack_keys(s, gid).astype(np.int64))
        gathered = [torch.empty_like(keys) for _ in range(world)]
        dist.all_gather(gathered, keys)
        merged2 = keyfmt.merge_keys(torch.stack(gathered).numpy().view(np.uint64), k)
        ms2, mi2 = keyfmt.unpack_keys(merged2)
        ss.compare_topk(ms2, mi2, truth, k)
        sizes = [None] * world
        dist.all_gather_object(sizes, len(mine))
        assert sum(sizes) == len(bank)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_sharded_search_protocol(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


def test_key_format_roundtrip_and_order():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import keyfmt
    s = np.array([[1.0, 0.5, 0.5, 0.0, -0.0, -1.0, np.inf, -np.inf, np.nan]], np.float32)
    i = np.arange(9, dtype=np.int64)[None]
    k = keyfmt.pack_keys(s, i)
    s2, i2 = keyfmt.unpack_keys(k)
    np.testing.assert_array_equal(i2, i)
    np.testing.assert_array_equal(np.isnan(s2), np.isnan(s))
    np.testing.assert_array_equal(s2[~np.isnan(s)], s[~np.isnan(s)])
    order = np.argsort(k[0])[::-1]
    # NaN first (torch.topk semantics), then +inf, 1, the 0.5 tie by ascending id, +0 before -0, ...
    assert order.tolist() == [8, 6, 0, 1, 2, 3, 4, 5, 7]
