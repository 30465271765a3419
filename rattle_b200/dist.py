"""Multi-GPU host logic (one process per GPU, torch.distributed for the plumbing).

clustering: every rank holds the whole read set; rank r extracts the k-mers of its block of reads and the blocks are
            exchanged (make_broadcast_callback); every rank then evaluates the (seed,target) pairs of each
            greedy wave whose target index t has t % world == rank, and the per-wave decision arrays (uint32, smaller
            wins, 0xffffffff = none) are min-reduced across ranks — the one real exchange step of the path.
correction: independent clusters -> round-robin shards, no collective.
"""
import numpy as np


def allreduce_min_u32_(t, group=None):
    """In-place unsigned-min all-reduce of an int32 tensor that holds uint32 bit patterns.
    x ^ 0x80000000 maps uint32 order onto int32 order, so the backend's signed MIN does the unsigned reduction."""
    import torch.distributed as dist
    t.bitwise_xor_(-2147483648)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    t.bitwise_xor_(-2147483648)
    return t


class _DevView:
    """zero-copy view of a raw device pointer (the decision arrays live in librattle_b200's own allocations)"""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 3}


def make_allreduce_callback(cuda_stream, group=None):
    """callback(ptr, count) for Context.set_shard: min-reduces `count` uint32 at device pointer `ptr` over NCCL.

    `cuda_stream` is the raw handle of the stream the Context runs on (the one passed to Context.set_stream; it must
    not be 0: the legacy default stream does not order against the library's own non-blocking stream).  The
    reduction is enqueued on that stream, between the kernels that produce and consume the decision arrays."""
    import torch
    if not cuda_stream:
        raise ValueError("make_allreduce_callback needs the explicit (non-default) CUDA stream the Context runs on")
    ext = torch.cuda.ExternalStream(cuda_stream)

    def cb(ptr, count):
        with torch.cuda.stream(ext):
            t = torch.as_tensor(_DevView(ptr, count), device="cuda")
            allreduce_min_u32_(t, group)
        return 0
    return cb


class _DevBytes:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (ptr, False), "version": 3}


def make_broadcast_callback(cuda_stream, group=None):
    """callback(ptr, nbytes, root) for Context.set_broadcast: the k-mer lists / bitvectors of rank `root`'s block of reads
    go to every rank over NCCL (sharded extraction), enqueued on the Context's stream."""
    import torch
    import torch.distributed as dist
    if not cuda_stream:
        raise ValueError("make_broadcast_callback needs the explicit (non-default) CUDA stream the Context runs on")
    ext = torch.cuda.ExternalStream(cuda_stream)
    chunk = 1 << 30  # (views stay below 2^31 elements)

    def cb(ptr, nbytes, root):
        with torch.cuda.stream(ext):
            at = 0
            while at < nbytes:
                n = min(chunk, nbytes - at)
                t = torch.as_tensor(_DevBytes(ptr + at, n), device="cuda")
                dist.broadcast(t, src=root, group=group)
                at += n
        return 0
    return cb


def shard_clusters(cl, rank, world):
    """clusters rank, rank+world, ... of a ClusterSet, with their global cluster ids"""
    from .api import ClusterSet
    nc = cl.n_clusters
    keep = (np.arange(nc) % world) == rank
    sizes = np.diff(cl.cl_off)
    off = np.zeros(int(keep.sum()) + 1, np.int64)
    off[1:] = np.cumsum(sizes[keep])
    mask = np.repeat(keep, sizes)
    total = int(cl.cl_off[-1])
    sub = ClusterSet(cl.main_id[keep].copy(), cl.main_rev[keep].copy(), off, cl.mem_id[:total][mask].copy(),
                     cl.mem_rev[:total][mask].copy(),
                     None if cl.main_gene is None else cl.main_gene[keep].copy(),
                     None if cl.mem_gene is None else cl.mem_gene[:total][mask].copy())
    return sub, np.nonzero(keep)[0]


def fastq_records(text: bytes):
    """4-line FASTQ records of a text as a list of bytes (without the trailing newline)"""
    lines = bytes(text).split(b"\n")
    return [b"\n".join(lines[i:i + 4]) for i in range(0, len(lines) - 1, 4)]


def merge_consensi(parts):
    """consensi.fq of the whole cluster set from the ranks' consensi texts: records in cluster-id order, which is the
    order correct_reads emits them in (correct.cpp:488-556 walks the clusters in clusters.out order).  The id is the
    one in the header ('@gene_cluster_<cid> ...' / '@transcript_cluster_<cid> ...'), i.e. the global id passed through
    rtl_set_cluster_ids."""
    recs = []
    for p in parts:
        for r in fastq_records(p):
            head = r.split(b" ", 1)[0]
            recs.append((int(head.rsplit(b"_", 1)[1]), r))
    recs.sort(key=lambda x: x[0])
    return b"".join(r + b"\n" for _, r in recs)
