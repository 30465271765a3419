"""ctypes binding of include/rattle_b200.h and the reference-shaped Python entry points."""
import ctypes
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "librattle_b200.so")

c_p = ctypes.c_void_p
c_i = ctypes.c_int
c_d = ctypes.c_double
c_i64 = ctypes.c_int64
c_u32 = ctypes.c_uint32


class RattleError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("rattle_b200 error %d: %s" % (code, msg))
        self.code = code


class Stats(ctypes.Structure):
    _fields_ = [("bv_pairs", c_i64), ("bv_launches", c_i64), ("bv_ms", c_d), ("full_pairs", c_i64),
                ("heavy_pairs", c_i64), ("join_ms", c_d), ("heavy_ms", c_d), ("extract_ms", c_d), ("waves", c_i64),
                ("rounds", c_i64), ("kernel_launches", c_i64), ("h2d_bytes", c_i64), ("d2h_bytes", c_i64),
                ("poa_alignments", c_i64), ("poa_cells", c_i64), ("poa_ms", c_d), ("poa_launches", c_i64),
                ("total_ms", c_d), ("poa_wall_ms", c_d), ("poa_busy_ms", c_d), ("upload_ms", c_d), ("poa_dram_bytes", c_i64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


ALLREDUCE_FN = ctypes.CFUNCTYPE(c_i, c_p, c_p, c_i64)
BROADCAST_FN = ctypes.CFUNCTYPE(c_i, c_p, c_p, c_i64, c_i)

# every symbol include/rattle_b200.h declares (tests/test_boundary.py checks the header against this list)
SYMBOLS = ["rtl_init", "rtl_destroy", "rtl_last_error", "rtl_set_option", "rtl_set_stream", "rtl_get_stats", "rtl_cluster_reads",
           "rtl_cluster_reads_batched", "rtl_sort_reads_by_length",
           "rtl_reads_upload", "rtl_cluster_resident", "rtl_set_shard", "rtl_set_broadcast", "rtl_extract_kmers", "rtl_bv_scan",
           "rtl_pair_similarity", "rtl_poa_msa", "rtl_correct_reads", "rtl_set_labels", "rtl_set_cluster_ids", "rtl_hps_encode", "rtl_hps_decode"]

_lib = None


def lib_path() -> str:
    return _LIB_PATH


def load_library():
    """dlopen librattle_b200.so. Raises (never falls back) when the CUDA library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise RattleError(-1, "librattle_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                              "there is no CPU fallback")
    L = ctypes.CDLL(_LIB_PATH)
    L.rtl_init.restype = c_i
    L.rtl_init.argtypes = [c_i, ctypes.POINTER(c_p)]
    L.rtl_destroy.restype = None
    L.rtl_destroy.argtypes = [c_p]
    L.rtl_last_error.restype = ctypes.c_char_p
    L.rtl_last_error.argtypes = [c_p]
    L.rtl_set_option.restype = c_i
    L.rtl_set_option.argtypes = [c_p, ctypes.c_char_p, c_i64]
    L.rtl_set_stream.restype = c_i
    L.rtl_set_stream.argtypes = [c_p, c_p]
    L.rtl_get_stats.restype = c_i
    L.rtl_get_stats.argtypes = [c_p, ctypes.POINTER(Stats)]
    L.rtl_cluster_reads.restype = c_i
    L.rtl_cluster_reads.argtypes = [c_p, c_p, c_p, c_u32, c_i, c_d, c_d, c_d, c_d, c_d, c_d, c_i, c_p, c_p, c_p, c_p,
                                    c_p, c_p]
    L.rtl_cluster_reads_batched.restype = c_i
    L.rtl_cluster_reads_batched.argtypes = [c_p, c_p, c_p, c_u32, c_p, c_u32, c_i, c_d, c_d, c_d, c_d, c_d, c_d, c_i, c_p, c_p,
                                            c_p, c_p, c_p, c_p, c_p]
    L.rtl_sort_reads_by_length.restype = c_i
    L.rtl_sort_reads_by_length.argtypes = [c_p, c_p, c_u32, c_p]
    L.rtl_reads_upload.restype = c_i
    L.rtl_reads_upload.argtypes = [c_p, c_p, c_p, c_u32]
    L.rtl_cluster_resident.restype = c_i
    L.rtl_cluster_resident.argtypes = [c_p, c_i, c_d, c_d, c_d, c_d, c_d, c_d, c_i, c_p, c_p, c_p, c_p, c_p, c_p]
    L.rtl_set_shard.restype = c_i
    L.rtl_set_shard.argtypes = [c_p, c_i, c_i, ALLREDUCE_FN, c_p]
    L.rtl_set_broadcast.restype = c_i
    L.rtl_set_broadcast.argtypes = [c_p, BROADCAST_FN, c_p]
    L.rtl_extract_kmers.restype = c_i
    L.rtl_extract_kmers.argtypes = [c_p, c_p, c_p, c_u32, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_p]
    L.rtl_set_labels.restype = c_i
    L.rtl_set_labels.argtypes = [c_p, c_p, c_i]
    L.rtl_set_cluster_ids.restype = c_i
    L.rtl_set_cluster_ids.argtypes = [c_p, c_p, c_i]
    L.rtl_bv_scan.restype = c_i
    L.rtl_bv_scan.argtypes = [c_p, c_i, c_i, c_p, c_i, c_p, c_i, c_d, c_p, c_p]
    L.rtl_pair_similarity.restype = c_i
    L.rtl_pair_similarity.argtypes = [c_p, c_i, c_i, c_p, c_p, c_p, c_i64, c_d, c_d, c_p, c_p, c_p, c_p, c_p]
    L.rtl_poa_msa.restype = c_i
    L.rtl_poa_msa.argtypes = [c_p, c_p, c_p, c_u32, c_i, c_i, c_i, c_i, c_p, c_i64, c_p, c_p, c_p, c_i64]
    L.rtl_correct_reads.restype = c_i
    L.rtl_correct_reads.argtypes = [c_p, c_p, c_p, c_p, c_u32, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_d,
                                    c_d, c_d, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_p]
    L.rtl_hps_encode.restype = c_i64
    L.rtl_hps_encode.argtypes = [c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i64]
    L.rtl_hps_decode.restype = c_i
    L.rtl_hps_decode.argtypes = [c_p, c_i64, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p]
    _lib = L
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(c_p)


def _bases(a) -> np.ndarray:
    if isinstance(a, (bytes, bytearray)):
        a = np.frombuffer(bytes(a), dtype=np.uint8)
    return np.ascontiguousarray(a, dtype=np.uint8)


@dataclass
class ClusterSet:
    """Flat cluster_set_t (cluster.hpp:10-42): cluster c = members mem_id/mem_rev[cl_off[c]:cl_off[c+1]]."""
    main_id: np.ndarray
    main_rev: np.ndarray
    cl_off: np.ndarray
    mem_id: np.ndarray
    mem_rev: np.ndarray
    main_gene: Optional[np.ndarray] = None
    mem_gene: Optional[np.ndarray] = None

    @property
    def n_clusters(self) -> int:
        return len(self.main_id)

    def as_dict(self):
        return dict(n_clusters=self.n_clusters, main_id=self.main_id, main_rev=self.main_rev, cl_off=self.cl_off,
                    mem_id=self.mem_id, mem_rev=self.mem_rev)


class Context:
    """One rtl_ctx (one GPU). Not thread-safe: serialise calls per context."""

    def __init__(self, device: int = 0):
        self.L = load_library()
        h = c_p()
        rc = self.L.rtl_init(device, ctypes.byref(h))
        if rc != 0:
            raise RattleError(rc, (self.L.rtl_last_error(None) or b"").decode())
        self.h = h
        self._cb = None

    def close(self):
        if getattr(self, "h", None):
            self.L.rtl_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise RattleError(rc, (self.L.rtl_last_error(self.h) or b"").decode())
        return rc

    def set_option(self, key: str, value: int):
        self._check(self.L.rtl_set_option(self.h, key.encode(), int(value)))

    def set_stream(self, cuda_stream: int):
        """cuda_stream: a cudaStream_t as int (e.g. torch.cuda.current_stream().cuda_stream); 0 = own stream."""
        self._check(self.L.rtl_set_stream(self.h, c_p(cuda_stream) if cuda_stream else None))

    def set_labels(self, labels):
        """file labels of `rattle correct -l` (consensus headers then carry per-label read counts)"""
        arr = (ctypes.c_char_p * max(1, len(labels)))(*[x.encode() if isinstance(x, str) else x for x in labels])
        self._check(self.L.rtl_set_labels(self.h, arr, len(labels)))

    def set_cluster_ids(self, ids):
        """global ids of the clusters passed to the following correct_reads calls (sharded correction); None resets"""
        if ids is None or len(ids) == 0:
            self._check(self.L.rtl_set_cluster_ids(self.h, None, 0))
        else:
            a = np.ascontiguousarray(ids, dtype=np.int32)
            self._check(self.L.rtl_set_cluster_ids(self.h, _ptr(a), len(a)))

    def stats(self) -> dict:
        s = Stats()
        self._check(self.L.rtl_get_stats(self.h, ctypes.byref(s)))
        return s.as_dict()

    def set_shard(self, rank: int, world: int, allreduce_min=None):
        """allreduce_min(device_ptr:int, count:int) -> 0 on success; min-reduces count uint32 in place across ranks."""
        if allreduce_min is None:
            cb = ctypes.cast(None, ALLREDUCE_FN)
        else:
            def _tramp(_user, ptr, count):
                try:
                    return int(allreduce_min(int(ptr), int(count)) or 0)
                except Exception:  # never let an exception cross the C ABI
                    import traceback
                    traceback.print_exc()
                    return -1
            cb = ALLREDUCE_FN(_tramp)
        self._cb = cb
        self._check(self.L.rtl_set_shard(self.h, rank, world, cb, None))

    def set_broadcast(self, broadcast=None):
        """broadcast(device_ptr:int, nbytes:int, root:int) -> 0 on success: copies nbytes at device_ptr from rank `root` to
        every other rank (sharded k-mer extraction, rtl_set_broadcast); None = every rank extracts every read."""
        if broadcast is None:
            cb = ctypes.cast(None, BROADCAST_FN)
        else:
            def _tramp(_user, ptr, nbytes, root):
                try:
                    return int(broadcast(int(ptr), int(nbytes), int(root)) or 0)
                except Exception:  # never let an exception cross the C ABI
                    import traceback
                    traceback.print_exc()
                    return -1
            cb = BROADCAST_FN(_tramp)
        self._bcb = cb
        self._check(self.L.rtl_set_broadcast(self.h, cb, None))

    # ---- hot path A
    def _cluster_out(self, n):
        return (np.zeros(n, np.int32), np.zeros(n, np.uint8), np.zeros(n + 1, np.int64), np.zeros(n, np.int32),
                np.zeros(n, np.uint8))

    @staticmethod
    def _pack(nc, out) -> ClusterSet:
        main_id, main_rev, cl_off, mem_id, mem_rev = out
        return ClusterSet(main_id[:nc].copy(), main_rev[:nc].copy(), cl_off[:nc + 1].copy(), mem_id, mem_rev)

    def cluster_reads(self, bases, offsets, kmer_size=10, t_s=0.2, t_v=1e6, bv_threshold=0.4, min_bv_threshold=0.2,
                      bv_falloff=0.05, repr_percentile=0.15, is_rna=False) -> ClusterSet:
        bases = _bases(bases)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        out = self._cluster_out(n)
        nc = ctypes.c_int32(0)
        self._check(self.L.rtl_cluster_reads(self.h, _ptr(bases), _ptr(offsets), n, kmer_size, t_s, t_v, bv_threshold,
                                             min_bv_threshold, bv_falloff, repr_percentile, int(is_rna),
                                             *[_ptr(a) for a in out], ctypes.byref(nc)))
        return self._pack(nc.value, out)

    def cluster_reads_batched(self, bases, offsets, seg_off, kmer_size=10, t_s=0.2, t_v=1e6, bv_threshold=0.4,
                              min_bv_threshold=0.2, bv_falloff=0.05, repr_percentile=0.15, is_rna=False):
        """n_seg independent cluster_reads problems in one pass (the per-gene loop of `rattle cluster --iso`,
        main.cpp:281-324): reads seg_off[s]:seg_off[s+1] are segment s.  Returns (ClusterSet, seg_cl_off): the clusters
        of segment s are seg_cl_off[s]:seg_cl_off[s+1], their ids count from the segment's first read."""
        bases = _bases(bases)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        seg_off = np.ascontiguousarray(seg_off, dtype=np.uint32)
        n = len(offsets) - 1
        n_seg = len(seg_off) - 1
        out = self._cluster_out(n)
        seg_cl_off = np.zeros(n_seg + 1, dtype=np.int64)
        nc = ctypes.c_int32(0)
        self._check(self.L.rtl_cluster_reads_batched(self.h, _ptr(bases), _ptr(offsets), n, _ptr(seg_off), n_seg, kmer_size,
                                                     t_s, t_v, bv_threshold, min_bv_threshold, bv_falloff, repr_percentile,
                                                     int(is_rna), *[_ptr(a) for a in out], ctypes.byref(nc), _ptr(seg_cl_off)))
        return self._pack(nc.value, out), seg_cl_off

    def sort_by_length(self, offsets) -> np.ndarray:
        """sort_read_set (fasta.cpp:458-464) on the GPU: the permutation that puts the reads into visitation order"""
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        perm = np.zeros(n, np.uint32)
        self._check(self.L.rtl_sort_reads_by_length(self.h, _ptr(offsets), n, _ptr(perm)))
        return perm

    def upload(self, bases, offsets):
        bases = _bases(bases)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self._n = len(offsets) - 1
        self._check(self.L.rtl_reads_upload(self.h, _ptr(bases), _ptr(offsets), self._n))

    def cluster_resident(self, kmer_size=10, t_s=0.2, t_v=1e6, bv_threshold=0.4, min_bv_threshold=0.2,
                         bv_falloff=0.05, repr_percentile=0.15, is_rna=False) -> ClusterSet:
        out = self._cluster_out(self._n)
        nc = ctypes.c_int32(0)
        self._check(self.L.rtl_cluster_resident(self.h, kmer_size, t_s, t_v, bv_threshold, min_bv_threshold,
                                                bv_falloff, repr_percentile, int(is_rna), *[_ptr(a) for a in out],
                                                ctypes.byref(nc)))
        return self._pack(nc.value, out)

    def extract_kmers(self, bases, offsets, kmer_size, both_strands=True):
        """Returns (fwd_hash, fwd_pos, rev_hash, rev_pos, bv_fwd[n,64], bv_rev[n,64]); list of read i starts at
        offsets[i]-i*k and has len_i-k entries."""
        bases = _bases(bases)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        total_k = int(offsets[-1] - offsets[0]) - kmer_size * n
        fh = np.zeros(total_k, np.uint32); fp = np.zeros(total_k, np.int32)
        rh = np.zeros(total_k, np.uint32); rp = np.zeros(total_k, np.int32)
        bf = np.zeros((n, 64), np.uint64); br = np.zeros((n, 64), np.uint64)
        self._n = n
        self._check(self.L.rtl_extract_kmers(self.h, _ptr(bases), _ptr(offsets), n, kmer_size, int(both_strands),
                                             _ptr(fh), _ptr(fp), _ptr(rh), _ptr(rp), _ptr(bf), _ptr(br)))
        return fh, fp, rh, rp, bf, br

    def bv_scan(self, seed_reads, target_reads, bv_threshold, kmer_size=10, is_rna=False, want_output=True):
        s = np.ascontiguousarray(seed_reads, np.int32)
        t = np.ascontiguousarray(target_reads, np.int32)
        if not want_output:  # kernel timing only (tools/bv_stream_bench.py): results stay on the device
            self._check(self.L.rtl_bv_scan(self.h, kmer_size, int(is_rna), _ptr(s), len(s), _ptr(t), len(t),
                                           bv_threshold, None, None))
            return None
        common = np.zeros((len(s), len(t)), np.uint32)
        passed = np.zeros((len(s), len(t)), np.uint8)
        self._check(self.L.rtl_bv_scan(self.h, kmer_size, int(is_rna), _ptr(s), len(s), _ptr(t), len(t), bv_threshold,
                                       _ptr(common), _ptr(passed)))
        return common & 0xffff, common >> 16, passed

    def pair_similarity(self, a_read, b_read, strand, kmer_size=10, is_rna=False, t_s=0.2, t_v=1e6):
        a = np.ascontiguousarray(a_read, np.int32)
        b = np.ascontiguousarray(b_read, np.int32)
        st = np.ascontiguousarray(strand, np.uint8)
        n = len(a)
        n_common = np.zeros(n, np.int64); bases = np.zeros(n, np.int32); nd = np.zeros(n, np.int32)
        var = np.zeros(n, np.float64); acc = np.zeros(n, np.uint8)
        self._check(self.L.rtl_pair_similarity(self.h, kmer_size, int(is_rna), _ptr(a), _ptr(b), _ptr(st), n, t_s, t_v,
                                               _ptr(n_common), _ptr(bases), _ptr(nd), _ptr(var), _ptr(acc)))
        return dict(n_common=n_common, bases=bases, n_dist=nd, var=var, accept=acc)

    # ---- hot path B
    def poa_msa(self, bases, offsets, m=5, n=-4, g=-8, e=-6, want_alignments=False):
        bases = _bases(bases)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        nseq = len(offsets) - 1
        total = int(offsets[-1] - offsets[0])
        cap = max(1, nseq) * (total + 16)
        msa = np.zeros(cap, np.uint8)
        cols = c_i(0)
        aln_off = np.zeros(nseq + 1, np.int64) if want_alignments else None
        aln_cap = 4 * (total + 16) if want_alignments else 0
        aln = np.zeros(max(2 * aln_cap, 2), np.int32) if want_alignments else None
        r = self._check(self.L.rtl_poa_msa(self.h, _ptr(bases), _ptr(offsets), nseq, m, n, g, e, _ptr(msa), cap,
                                           ctypes.byref(cols), _ptr(aln_off), _ptr(aln), aln_cap))
        rows = [msa[i * cols.value:(i + 1) * cols.value].tobytes() for i in range(r)]
        if want_alignments:
            return rows, [aln[2 * aln_off[i]:2 * aln_off[i + 1]].reshape(-1, 2).copy() for i in range(nseq)]
        return rows

    def correct_reads(self, bases, quals, offsets, clusters: ClusterSet, min_occ=0.3, gap_occ=0.3, err_ratio=30.0,
                      split=200, min_reads=5, headers=None, as_bytes=True, cluster_ids=None):
        """correct_reads (correct.hpp:44): returns (corrected, uncorrected, consensi) FASTQ text as bytes.

        as_bytes=False returns uint8 views of the Context's (reused) output buffers instead: they are what the library
        wrote into caller-owned host memory, valid until the next correct_reads call on this Context — no extra copy
        into Python bytes objects and no fresh pages to fault in on every call (bench.py's timed loop)."""
        bases = _bases(bases)
        quals = _bases(quals)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        nc = clusters.n_clusters
        # the C side reads these through raw pointers: enforce the dtypes of include/rattle_b200.h
        clusters = ClusterSet(np.ascontiguousarray(clusters.main_id, np.int32), np.ascontiguousarray(clusters.main_rev, np.uint8),
                              np.ascontiguousarray(clusters.cl_off, np.int64), np.ascontiguousarray(clusters.mem_id, np.int32),
                              np.ascontiguousarray(clusters.mem_rev, np.uint8), clusters.main_gene, clusters.mem_gene)
        self.set_cluster_ids(cluster_ids)
        gm = np.full(nc, -1, np.int32) if clusters.main_gene is None else np.ascontiguousarray(clusters.main_gene, np.int32)
        gs = (np.full(len(clusters.mem_id), -1, np.int32) if clusters.mem_gene is None
              else np.ascontiguousarray(clusters.mem_gene, np.int32))
        hdr = hoff = None
        if headers is not None:
            hoff = np.zeros(n + 1, np.uint64)
            hoff[1:] = np.cumsum([len(h) for h in headers])
            hdr = np.frombuffer(b"".join(headers), dtype=np.uint8).copy()
        cap = 4 * int(offsets[-1]) + 256 * (n + nc) + 1024
        while True:
            held = getattr(self, "_corr_bufs", None)
            if held is None or len(held[0]) < cap:
                held = self._corr_bufs = [np.empty(cap, np.uint8) for _ in range(3)]  # untouched pages cost nothing
            bufs = held
            cap = len(bufs[0])
            lens = [c_i64(cap) for _ in range(3)]
            args = [self.h, _ptr(bases), _ptr(quals), _ptr(offsets), n, _ptr(hdr), _ptr(hoff), _ptr(clusters.main_id),
                    _ptr(clusters.main_rev), _ptr(gm), _ptr(clusters.cl_off), _ptr(clusters.mem_id),
                    _ptr(clusters.mem_rev), _ptr(gs), nc, min_occ, gap_occ, err_ratio, split, min_reads]
            for b, l in zip(bufs, lens):
                args += [_ptr(b), ctypes.byref(l)]
            rc = self.L.rtl_correct_reads(*args)
            if rc == -3 and max(l.value for l in lens) > cap:
                cap = max(l.value for l in lens) + 1024
                continue
            self._check(rc)
            views = tuple(b[:l.value] for b, l in zip(bufs, lens))
            return tuple(v.tobytes() for v in views) if as_bytes else views


def hps_encode(cl: ClusterSet) -> bytes:
    """clusters.out bytes (main.cpp:275)."""
    L = load_library()
    cap = 16 + 12 * (len(cl.mem_id) + 2 * cl.n_clusters)
    out = np.zeros(cap, np.uint8)
    i32 = lambda a: None if a is None else np.ascontiguousarray(a, np.int32)  # noqa: E731
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)  # noqa: E731
    main_id, mem_id, main_gene, mem_gene = i32(cl.main_id), i32(cl.mem_id), i32(cl.main_gene), i32(cl.mem_gene)
    main_rev, mem_rev, cl_off = u8(cl.main_rev), u8(cl.mem_rev), np.ascontiguousarray(cl.cl_off, np.int64)
    n = L.rtl_hps_encode(cl.n_clusters, _ptr(main_id), _ptr(main_rev), _ptr(main_gene), _ptr(cl_off),
                         _ptr(mem_id), _ptr(mem_rev), _ptr(mem_gene), _ptr(out), cap)
    if n < 0:
        raise RattleError(-3, "hps buffer too small")
    return out[:n].tobytes()


def hps_decode(buf: bytes) -> ClusterSet:
    L = load_library()
    b = np.frombuffer(buf, dtype=np.uint8)
    nc = ctypes.c_int32(0)
    nm = c_i64(0)
    rc = L.rtl_hps_decode(_ptr(b), len(b), ctypes.byref(nc), ctypes.byref(nm), None, None, None, None, None, None, None)
    if rc != 0:
        raise RattleError(rc, "malformed clusters.out")
    cs = ClusterSet(np.zeros(nc.value, np.int32), np.zeros(nc.value, np.uint8), np.zeros(nc.value + 1, np.int64),
                    np.zeros(nm.value, np.int32), np.zeros(nm.value, np.uint8), np.zeros(nc.value, np.int32),
                    np.zeros(nm.value, np.int32))
    rc = L.rtl_hps_decode(_ptr(b), len(b), ctypes.byref(nc), ctypes.byref(nm), _ptr(cs.main_id), _ptr(cs.main_rev),
                          _ptr(cs.main_gene), _ptr(cs.cl_off), _ptr(cs.mem_id), _ptr(cs.mem_rev), _ptr(cs.mem_gene))
    if rc != 0:
        raise RattleError(rc, "malformed clusters.out")
    return cs


_default_ctx = None


def _ctx() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(int(os.environ.get("LOCAL_RANK", "0")))
    return _default_ctx


def cluster_reads(bases, offsets, kmer_size, t_s, t_v, bv_threshold, min_bv_threshold, bv_falloff,
                  min_reads_cluster=0, use_hc=False, repr_percentile=0.15, is_rna=False, verbose=False,
                  n_threads=1) -> ClusterSet:
    """Same argument list as the reference's cluster_reads (cluster.hpp:44); `min_reads_cluster`, `use_hc`
    (ignored / always false in the reference) and `n_threads` are accepted and unused."""
    return _ctx().cluster_reads(bases, offsets, kmer_size, t_s, t_v, bv_threshold, min_bv_threshold, bv_falloff,
                                repr_percentile, is_rna)


def correct_reads(clusters: ClusterSet, bases, quals, offsets, min_occ, gap_occ, err_ratio, split, min_reads,
                  n_threads=1, verbose=False, labels=None, headers=None):
    """Same argument meaning as the reference's correct_reads (correct.hpp:44)."""
    _ctx().set_labels(list(labels or []))
    return _ctx().correct_reads(bases, quals, offsets, clusters, min_occ, gap_occ, err_ratio, split, min_reads,
                                headers=headers)
