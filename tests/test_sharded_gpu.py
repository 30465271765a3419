"""The N>1 path on ONE GPU: rank/world sharding of the clustering kernels (rtl_set_shard) with the exchange step done by a
device-side min-reduction between several contexts running in threads, and sharded correction with global cluster
ids (rtl_set_cluster_ids).  Bar: bit-identical to the unsharded run (and to the oracle where it is cheap).

This is what bench.py does over NCCL with one process per GPU; here the ranks are threads of one process so that the
test runs on the driver's 1-GPU box.  The reduction callback contract is the C ABI's (include/rattle_b200.h):
min-reduce `count` uint32 at `device_ptr` across ranks, in place, ordered against the context's stream.
"""
import threading

import numpy as np
import pytest

from tools import synth

pytestmark = pytest.mark.gpu


class _DevView:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 3}


class _ByteView:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (ptr, False), "version": 3}


class ThreadRanks:
    """world contexts on device 0, one thread each; the callback min-reduces the ranks' buffers on the device"""

    def __init__(self, world, options=None, broadcast=True):
        import torch
        import rattle_b200
        self.torch = torch
        self.world = world
        self.barrier = threading.Barrier(world)
        self.views = [None] * world
        self.streams = [torch.cuda.Stream() for _ in range(world)]
        self.ctxs = [rattle_b200.Context(0) for _ in range(world)]
        self.calls = 0
        self.bviews = [None] * world
        self.bcasts = 0
        for r, c in enumerate(self.ctxs):
            for k, v in (options or {}).items():
                c.set_option(k, v)
            c.set_stream(self.streams[r].cuda_stream)
            c.set_shard(r, world, self._callback(r))
            if broadcast:  # sharded k-mer extraction (rtl_set_broadcast); without it every rank extracts every read
                c.set_broadcast(self._broadcast(r))

    def _callback(self, rank):
        torch = self.torch

        def cb(ptr, count):
            self.streams[rank].synchronize()  # this rank's producers are done
            self.views[rank] = torch.as_tensor(_DevView(ptr, count), device="cuda")
            self.barrier.wait()
            if rank == 0:
                assert len({v.numel() for v in self.views}) == 1, "ranks disagree on the exchange size"
                flip = -2147483648  # uint32 order == int32 order after flipping the sign bit
                m = self.views[0] ^ flip
                for v in self.views[1:]:
                    m = torch.minimum(m, v ^ flip)
                m ^= flip
                for v in self.views:
                    v.copy_(m)
                torch.cuda.synchronize()
                self.calls += 1
            self.barrier.wait()
            return 0
        return cb

    def _broadcast(self, rank):
        """rtl_set_broadcast's contract between threads: rank `root`'s bytes go to the same address range of the others"""
        torch = self.torch

        def cb(ptr, nbytes, root):
            self.streams[rank].synchronize()
            self.bviews[rank] = torch.as_tensor(_ByteView(ptr, nbytes), device="cuda")
            self.barrier.wait()
            if rank == 0:
                assert len({v.numel() for v in self.bviews}) == 1, "ranks disagree on the broadcast size"
                for r, v in enumerate(self.bviews):
                    if r != root:
                        v.copy_(self.bviews[root])
                torch.cuda.synchronize()
                self.bcasts += 1
            self.barrier.wait()
            return 0
        return cb

    def run(self, fn):
        out, err = [None] * self.world, [None] * self.world

        def work(r):
            try:
                out[r] = fn(r, self.ctxs[r])
            except BaseException as e:  # a dead rank would leave the others at the barrier
                err[r] = e
                self.barrier.abort()
        th = [threading.Thread(target=work, args=(r,)) for r in range(self.world)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        for e in err:
            if e is not None and not isinstance(e, threading.BrokenBarrierError):
                raise e
        for e in err:
            if e is not None:
                raise e
        return out

    def close(self):
        for c in self.ctxs:
            c.close()


def same(a, b):
    return all(np.array_equal(getattr(a, k), getattr(b, k)) for k in ("main_id", "main_rev", "cl_off", "mem_id", "mem_rev"))


@pytest.mark.parametrize("world,wave", [(2, 512), (3, 64)])
def test_sharded_clustering_equals_unsharded_and_oracle(ctx, orc, world, wave):
    """initial pass + merge rounds (with the memo of failed representative pairs) under rank/world sharding"""
    rs = synth.generate(seed=33, n_genes=40, reads_per_tx=25, len_mean=900.0, len_sd=120.0, len_min=400,
                        len_max=2000).sorted_by_length()[0]
    single = ctx.cluster_reads(rs.bases, rs.offsets, is_rna=False)
    exp = orc.cluster_reads(rs.bases, rs.offsets, is_rna=False, n_threads=8)
    assert single.n_clusters == exp["n_clusters"]
    ranks = ThreadRanks(world, {"wave": wave})
    try:
        outs = ranks.run(lambda r, c: c.cluster_reads(rs.bases, rs.offsets, is_rna=False))
        pairs = [c.stats()["bv_pairs"] for c in ranks.ctxs]
    finally:
        ranks.close()
    assert ranks.calls > 0 and ranks.bcasts > 0
    for o in outs:
        assert same(o, single)
        for k in ("main_id", "main_rev", "cl_off", "mem_id", "mem_rev"):
            assert np.array_equal(getattr(o, k), exp[k]), k
    # the ranks really split the work: nobody evaluated (nearly) all pairs
    total = ctx.stats()["bv_pairs"]
    assert max(pairs) < 0.8 * total, (pairs, total)


def test_sharded_extraction_exchanges_the_blocks(ctx):
    """rtl_set_broadcast: rank r extracts reads [r*n/W, (r+1)*n/W) only; after the exchange every rank holds every read's
    k-mer lists and bitvectors, bit-identical to an unsharded extraction; without the callback the ranks extract everything
    themselves (same result); a base outside ACGTU in ONE rank's block fails the call on EVERY rank"""
    import rattle_b200
    rs = synth.generate(seed=36, n_genes=11, reads_per_tx=9, len_mean=800.0, len_sd=200.0, len_min=300, len_max=1800).sorted_by_length()[0]
    want = ctx.extract_kmers(rs.bases, rs.offsets, 10, True)
    for bc in (True, False):
        ranks = ThreadRanks(3, broadcast=bc)
        try:
            outs = ranks.run(lambda r, c: c.extract_kmers(rs.bases, rs.offsets, 10, True))
        finally:
            ranks.close()
        assert (ranks.bcasts > 0) == bc
        for o in outs:
            for a, b in zip(o, want):
                assert np.array_equal(a, b)
    bad = rs.bases.copy()
    bad[int(rs.offsets[rs.n - 2]) + 5] = ord("N")  # in the last rank's block
    ranks = ThreadRanks(3)
    errs = []

    def attempt(r, c):
        try:
            c.extract_kmers(bad, rs.offsets, 10, True)
        except rattle_b200.RattleError as e:
            errs.append((r, e.code))
    try:
        ranks.run(attempt)
    finally:
        ranks.close()
    assert sorted(r for r, _ in errs) == [0, 1, 2] and all(code == -2 for _, code in errs)


def test_sharded_clustering_rna_single_strand(ctx):
    rs = synth.generate(seed=34, n_genes=25, reads_per_tx=20, len_mean=700.0, len_sd=80.0, len_min=300, len_max=1500,
                        p_flip=0.0).sorted_by_length()[0]
    single = ctx.cluster_reads(rs.bases, rs.offsets, is_rna=True)
    ranks = ThreadRanks(2)
    try:
        outs = ranks.run(lambda r, c: c.cluster_reads(rs.bases, rs.offsets, is_rna=True))
    finally:
        ranks.close()
    assert all(same(o, single) for o in outs)


def _records(text):
    lines = text.split(b"\n")
    return [b"\n".join(lines[i:i + 4]) for i in range(0, len(lines) - 1, 4)]


def test_sharded_correction_carries_global_cluster_ids(ctx, ref):
    """clusters r, r+W, ... corrected per rank with rtl_set_cluster_ids: the union of the ranks' files equals the
    single-rank run and the reference (headers carry the index in the WHOLE cluster set, correct.cpp:344-349,540-549)"""
    from rattle_b200.dist import merge_consensi, shard_clusters
    rs = synth.generate(seed=35, n_genes=9, reads_per_tx=14, len_mean=500.0, len_sd=50.0, len_min=300,
                        len_max=900).sorted_by_length()[0]
    cl = ctx.cluster_reads(rs.bases, rs.offsets, is_rna=False)
    assert cl.n_clusters >= 4
    whole = ctx.correct_reads(rs.bases, rs.quals, rs.offsets, cl, min_reads=5)
    exp = ref.correct_reads(rs.bases, rs.quals, rs.offsets, cl.as_dict(), min_reads=5)
    assert whole[2] == exp[2] and whole[1] == exp[1]
    world = 3
    parts = []
    for r in range(world):
        sub, gids = shard_clusters(cl, r, world)
        parts.append(ctx.correct_reads(rs.bases, rs.quals, rs.offsets, sub, min_reads=5, cluster_ids=gids))
    ctx.set_cluster_ids(None)
    assert merge_consensi([p[2] for p in parts]) == whole[2]
    for k in (0, 1):
        assert sorted(sum((_records(p[k]) for p in parts), [])) == sorted(_records(whole[k]))
    # without the ids the shards would renumber their clusters from 0 (the bug ADVICE r01 describes)
    sub, gids = shard_clusters(cl, 1, world)
    local = ctx.correct_reads(rs.bases, rs.quals, rs.offsets, sub, min_reads=5)
    assert local[2] != parts[1][2] and b"@gene_cluster_0 " in local[2]
