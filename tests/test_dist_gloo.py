"""CPU, world_size 2, gloo: the host-side logic of the N>1 path (rattle_b200/dist.py) — the unsigned-min exchange of
per-wave decision arrays and the cluster sharding used for correction."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from rattle_b200 import ClusterSet
        from rattle_b200.dist import allreduce_min_u32_, shard_clusters
        # decision arrays: 0xffffffff = "no seed matched", otherwise 2*seed+strand; each rank saw only its targets
        rng = np.random.default_rng(5)
        full = rng.integers(0, 1000, 64).astype(np.uint32)
        full[rng.random(64) < 0.3] = 0xffffffff
        mine = full.copy()
        mine[np.arange(64) % world != rank] = 0xffffffff
        t = torch.from_numpy(mine.view(np.int32).copy())
        allreduce_min_u32_(t)
        ok_min = bool(np.array_equal(t.numpy().view(np.uint32), full))
        # a value >= 2^31 must still lose against a small one and win against the sentinel
        t2 = torch.from_numpy(np.array([0x80000005 if rank == 0 else 0xffffffff, 7 if rank == 1 else 0x90000000],
                                       np.uint32).view(np.int32).copy())
        allreduce_min_u32_(t2)
        ok_big = t2.numpy().view(np.uint32).tolist() == [0x80000005, 7]
        sizes = np.array([3, 1, 4, 1, 5, 9, 2])
        off = np.zeros(8, np.int64)
        off[1:] = np.cumsum(sizes)
        ids = np.arange(int(off[-1]), dtype=np.int32)
        cl = ClusterSet(ids[off[:-1]].copy(), np.zeros(7, np.uint8), off, ids, (ids % 2).astype(np.uint8))
        sub, gids = shard_clusters(cl, rank, world)
        got = [None] * world
        dist.all_gather_object(got, (gids.tolist(), sub.mem_id.tolist(), sub.cl_off.tolist()))
        q.put((rank, ok_min, ok_big, got))
    finally:
        dist.destroy_process_group()


def test_world2_exchange_and_sharding():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_min, ok_big, got in res:
        assert ok_min and ok_big
        all_ids = sorted(g for part in got for g in part[0])
        assert all_ids == list(range(7))  # every cluster on exactly one rank
        members = sorted(m for part in got for m in part[1])
        assert members == list(range(25))
        assert got[0][0] == [0, 2, 4, 6] and got[1][0] == [1, 3, 5]
        assert got[1][2] == [0, 1, 2, 11]


def test_allreduce_callback_refuses_the_default_stream():
    """The NCCL exchange must be enqueued on the stream the library runs on; handle 0 (the legacy default stream) would
    silently fall back to the library's own non-blocking stream, which the default stream does not order against
    (this raced on 2 x B200 before the callback took the stream explicitly)."""
    import pytest
    from rattle_b200.dist import make_allreduce_callback
    with pytest.raises(ValueError):
        make_allreduce_callback(0)
