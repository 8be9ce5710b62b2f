"""N > 1 host logic on CPU: world_size-2 (and 3, ragged) `gloo` runs of the read sharding + gather
that bench.py --gpus N uses (stralg_b200/shard.py).  The per-shard search is the ORACLE here (test
infrastructure standing in for the device kernel); what is checked is the partition, the padding
of ragged shards and the order of the gathered (L, R) pairs against one unsharded oracle search
(the reference's per-read loop, tools/readmappers/bwt_readmapper/bwt_readmapper.c:128-161)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from stralg_b200.shard import max_shard, shard_bounds  # noqa: E402


def test_shard_bounds_cover_everything_once():
    for total in (0, 1, 2, 7, 100, 1001, 10**8):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            sizes = []
            for r in range(world):
                lo, hi = shard_bounds(total, world, r)
                assert lo == prev and hi >= lo
                prev = hi
                sizes.append(hi - lo)
            assert prev == total
            assert max(sizes) - min(sizes) <= 1
            assert max(sizes) == max_shard(total, world) or total == 0
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, m, chunks, out_path):
    import torch
    import torch.distributed as dist
    from _oracle import Oracle
    from stralg_b200.shard import ShardedSearch
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        o = Oracle()
        n = 20_000
        codes = o.random_codes(n, 4, seed=7)
        sa = o.sa(codes)
        bwt = o.bwt(codes, sa)
        c = o.c_table(codes, 5)
        ck = o.o_checkpoints(bwt, 5, 64)
        rng = np.random.default_rng(11)  # same reads on every rank; each takes its shard
        starts = rng.integers(0, n - m, total)
        reads = np.concatenate([codes[s:s + m] for s in starts]) if total else np.zeros(0, np.uint8)
        miss = rng.random(total) < 0.2
        for q in np.nonzero(miss)[0]:
            reads[q * m:(q + 1) * m] = rng.integers(1, 5, m)

        ss = ShardedSearch(total, m, "cpu", dist, chunks=chunks)
        assert (ss.lo, ss.hi) == shard_bounds(total, world, rank)
        mine = torch.from_numpy(reads[ss.lo * m: ss.hi * m].copy())

        def search_fn(r, mm, count, Lo, Ro):
            off = np.arange(0, (count + 1) * mm, mm, dtype=np.uint64)
            L, R = o.search_ck(c, bwt, ck, 64, r.numpy(), off)
            Lo[:count] = torch.from_numpy(L.view(np.int32))
            Ro[:count] = torch.from_numpy(R.view(np.int32))

        for _ in range(2):  # a step is repeatable (buffers are reused)
            ss.step(search_fn, mine)
        res = ss.result()
        if rank == 0:
            L, R = res
            off = np.arange(0, (total + 1) * m, m, dtype=np.uint64)
            Le, Re = o.search_ck(c, bwt, ck, 64, reads, off)
            ok = np.array_equal(L.numpy().view(np.uint32), Le) and np.array_equal(R.numpy().view(np.uint32), Re)
            hits = int((Re > Le).sum())
            with open(out_path, "w") as f:
                f.write(f"{int(ok)} {hits}")
        else:
            assert res is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,total,chunks", [(2, 1000, 1), (2, 1001, 4), (3, 1000, 3), (2, 1, 2)])
def test_sharded_search_gather_gloo(tmp_path, world, total, chunks):
    import torch.multiprocessing as mp
    out = str(tmp_path / "res.txt")
    mp.spawn(_worker, args=(world, _free_port(), total, 20, chunks, out), nprocs=world, join=True)
    ok, hits = open(out).read().split()
    assert ok == "1"
    if total >= 1000:
        assert int(hits) >= 0.7 * total  # most reads were sampled from the text
