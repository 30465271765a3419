"""TEST INFRASTRUCTURE ONLY — ctypes loaders for liboracle.so (our CPU restatement) and, when it was built,
oracle/_ref/libref_shim.so (the unmodified reference behind an extern "C" shim).

Allowed importers: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline / --impl reference legs).
rattle_b200/ must never import this package (tests/test_boundary.py greps for it).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "liboracle.so")
REF_SHIM_SO = os.path.join(_HERE, "_ref", "libref_shim.so")
REF_RATTLE = os.path.join(_HERE, "_ref", "rattle")
REFERENCE_ROOT = os.environ.get("RATTLE_REFERENCE", "/root/reference")

c_p = ctypes.c_void_p
c_i = ctypes.c_int
c_d = ctypes.c_double
c_i64 = ctypes.c_int64
c_u32 = ctypes.c_uint32


def build(with_ref: bool = True) -> None:
    """Compile liboracle.so, and oracle/_ref from the reference sources when they are present."""
    target = ["all"] if (with_ref and os.path.exists(os.path.join(REFERENCE_ROOT, "cluster.cpp"))) else ["liboracle.so"]
    subprocess.check_call(["make", "-s", "-C", _HERE, "REF=" + REFERENCE_ROOT, "-j8"] + target)


def have_ref() -> bool:
    return os.path.exists(REF_SHIM_SO)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(c_p)


class _Lib:
    """Common front end: `prefix` is 'orc' (restatement) or 'ref' (reference shim); same call shapes."""

    def __init__(self, path, prefix):
        self.lib = ctypes.CDLL(path)
        self.p = prefix
        L = self.lib
        f = getattr(L, prefix + "_extract_kmers")
        f.restype = c_i
        f.argtypes = [c_p, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_p]
        f = getattr(L, prefix + "_common_kmers")
        f.restype = c_i64
        f.argtypes = [c_p, c_p, c_i, c_p, c_p, c_i, c_p, c_p, c_i64]
        f = getattr(L, prefix + "_similarity")
        f.restype = c_i
        f.argtypes = [c_p, c_p, c_i64, c_i, c_p, c_p, c_i]
        f = getattr(L, prefix + "_var")
        f.restype = c_d
        f.argtypes = [c_p, c_i]
        f = getattr(L, prefix + "_pair_match")
        f.restype = c_i
        f.argtypes = [c_p, c_i, c_p, c_i, c_i, c_d, c_d, c_d, c_i]
        f = getattr(L, prefix + "_cluster_reads")
        f.restype = c_i
        f.argtypes = [c_p, c_p, c_u32, c_i, c_d, c_d, c_d, c_d, c_d, c_d, c_i, c_i, c_p, c_p, c_p, c_p, c_p]
        f = getattr(L, prefix + "_poa_msa")
        f.restype = c_i
        f.argtypes = [c_p, c_p, c_u32, c_i, c_i, c_i, c_i, c_p, c_i64, c_p, c_p, c_p, c_i64]
        f = getattr(L, prefix + "_correct_reads")
        f.restype = c_i
        if prefix == "ref":
            f.argtypes = [c_p, c_p, c_p, c_u32, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_d, c_d, c_d, c_i, c_i, c_i,
                          c_p, c_p, c_p, c_p, c_p, c_p]
        else:
            f.argtypes = [c_p, c_p, c_p, c_u32, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_d, c_d, c_d, c_i, c_i,
                          c_p, c_p, c_p, c_p, c_p, c_p]
        if prefix == "orc":
            L.orc_cluster_stats.restype = None
            L.orc_cluster_stats.argtypes = [c_p]
            L.orc_hps_encode.restype = c_i64
            L.orc_hps_encode.argtypes = [c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i64]
            L.orc_poa_cells.restype = c_i64
            L.orc_poa_cells.argtypes = []

    # -- k-mers -------------------------------------------------------------------------------
    def extract_kmers(self, seq: bytes, k: int, both: bool):
        n = len(seq) - k
        fh = np.zeros(max(n, 0), np.uint32); fp = np.zeros(max(n, 0), np.int32)
        rh = np.zeros(max(n, 0), np.uint32); rp = np.zeros(max(n, 0), np.int32)
        bf = np.zeros(64, np.uint64); br = np.zeros(64, np.uint64)
        r = getattr(self.lib, self.p + "_extract_kmers")(seq, len(seq), k, int(both), _ptr(fh), _ptr(fp), _ptr(rh),
                                                          _ptr(rp), _ptr(bf), _ptr(br))
        return r, fh, fp, rh, rp, bf, br

    def common_kmers(self, h1, p1, h2, p2):
        cap = 1 << 16
        while True:
            a = np.zeros(cap, np.int32); b = np.zeros(cap, np.int32)
            n = getattr(self.lib, self.p + "_common_kmers")(_ptr(h1), _ptr(p1), len(h1), _ptr(h2), _ptr(p2), len(h2),
                                                            _ptr(a), _ptr(b), cap)
            if n <= cap:
                return a[:n].copy(), b[:n].copy()
            cap = int(n)

    def similarity(self, first, second, k):
        first = np.ascontiguousarray(first, np.int32); second = np.ascontiguousarray(second, np.int32)
        bases = c_i(0)
        d = np.zeros(max(len(first), 1), np.int32)
        nd = getattr(self.lib, self.p + "_similarity")(_ptr(first), _ptr(second), len(first), k, ctypes.byref(bases),
                                                       _ptr(d), len(d))
        return bases.value, d[:nd].copy()

    def var(self, d):
        d = np.ascontiguousarray(d, np.int32)
        return getattr(self.lib, self.p + "_var")(_ptr(d), len(d))

    def pair_match(self, s1: bytes, s2: bytes, k, t_s, t_v, thr, is_rna):
        return getattr(self.lib, self.p + "_pair_match")(s1, len(s1), s2, len(s2), k, t_s, t_v, thr, int(is_rna))

    # -- clustering ---------------------------------------------------------------------------
    def cluster_reads(self, bases, offsets, k=10, t_s=0.2, t_v=1e6, bv_thr=0.4, bv_min=0.2, bv_falloff=0.05,
                      repr_pct=0.15, is_rna=False, n_threads=1):
        n = len(offsets) - 1
        main_id = np.zeros(n, np.int32); main_rev = np.zeros(n, np.uint8)
        cl_off = np.zeros(n + 1, np.int64)
        mem_id = np.zeros(n, np.int32); mem_rev = np.zeros(n, np.uint8)
        nc = getattr(self.lib, self.p + "_cluster_reads")(_ptr(bases), _ptr(offsets), n, k, t_s, t_v, bv_thr, bv_min,
                                                          bv_falloff, repr_pct, int(is_rna), n_threads, _ptr(main_id),
                                                          _ptr(main_rev), _ptr(cl_off), _ptr(mem_id), _ptr(mem_rev))
        if nc < 0:
            raise RuntimeError("%s_cluster_reads failed: %d" % (self.p, nc))
        return dict(n_clusters=nc, main_id=main_id[:nc].copy(), main_rev=main_rev[:nc].copy(),
                    cl_off=cl_off[:nc + 1].copy(), mem_id=mem_id, mem_rev=mem_rev)

    def cluster_stats(self):
        out = np.zeros(4, np.int64)
        self.lib.orc_cluster_stats(_ptr(out))
        return dict(bv_tests=int(out[0]), full=int(out[1]), accepted=int(out[2]), rounds=int(out[3]))

    def hps_encode(self, cl, gene_main=None, gene_mem=None) -> bytes:
        cap = 16 + 12 * (len(cl["mem_id"]) + 2 * cl["n_clusters"])
        out = np.zeros(cap, np.uint8)
        n = self.lib.orc_hps_encode(cl["n_clusters"], _ptr(cl["main_id"]), _ptr(cl["main_rev"]), _ptr(gene_main),
                                    _ptr(cl["cl_off"]), _ptr(cl["mem_id"]), _ptr(cl["mem_rev"]), _ptr(gene_mem),
                                    _ptr(out), cap)
        assert n >= 0
        return out[:n].tobytes()

    # -- POA ----------------------------------------------------------------------------------
    def poa_msa(self, bases, offsets, m=5, n=-4, g=-8, e=-6, want_alignments=False):
        nseq = len(offsets) - 1
        total = int(offsets[-1])
        cap = max(1, nseq) * (total + 16)
        msa = np.zeros(cap, np.uint8)
        cols = c_i(0)
        aln_off = np.zeros(nseq + 1, np.int64) if want_alignments else None
        aln_cap = 4 * (total + 16) * 2 if want_alignments else 0
        aln = np.zeros(max(aln_cap, 1), np.int32) if want_alignments else None
        r = getattr(self.lib, self.p + "_poa_msa")(_ptr(bases), _ptr(offsets), nseq, m, n, g, e, _ptr(msa), cap,
                                                   ctypes.byref(cols), _ptr(aln_off), _ptr(aln), aln_cap)
        if r < 0:
            raise RuntimeError("%s_poa_msa failed" % self.p)
        rows = [msa[i * cols.value:(i + 1) * cols.value].tobytes() for i in range(r)]
        if want_alignments:
            alns = [aln[2 * aln_off[i]:2 * aln_off[i + 1]].reshape(-1, 2).copy() for i in range(nseq)]
            return rows, alns
        return rows

    def correct_reads(self, bases, quals, offsets, cl, gene_main=None, gene_mem=None, min_occ=0.3, gap_occ=0.3,
                      err_ratio=30.0, split=200, min_reads=5, n_threads=1):
        n = len(offsets) - 1
        nc = cl["n_clusters"]
        gm = np.full(nc, -1, np.int32) if gene_main is None else np.ascontiguousarray(gene_main, np.int32)
        gs = np.full(len(cl["mem_id"]), -1, np.int32) if gene_mem is None else np.ascontiguousarray(gene_mem, np.int32)
        cap = 4 * int(offsets[-1]) + 256 * (n + nc) + 1024
        bufs = [np.zeros(cap, np.uint8) for _ in range(3)]
        lens = [c_i64(cap) for _ in range(3)]
        args = [_ptr(bases), _ptr(quals), _ptr(offsets), n, _ptr(cl["main_id"]), _ptr(cl["main_rev"]), _ptr(gm),
                _ptr(cl["cl_off"]), _ptr(cl["mem_id"]), _ptr(cl["mem_rev"]), _ptr(gs), nc, min_occ, gap_occ, err_ratio,
                split, min_reads]
        if self.p == "ref":
            args.append(n_threads)
        for b, l in zip(bufs, lens):
            args += [_ptr(b), ctypes.byref(l)]
        r = getattr(self.lib, self.p + "_correct_reads")(*args)
        if r != 0:
            raise RuntimeError("%s_correct_reads failed: %d" % (self.p, r))
        return tuple(b[:l.value].tobytes() for b, l in zip(bufs, lens))

    def poa_cells(self):
        return int(self.lib.orc_poa_cells())


_orc = None
_ref = None


def oracle() -> _Lib:
    global _orc
    if _orc is None:
        if not os.path.exists(ORACLE_SO):
            build(with_ref=False)
        _orc = _Lib(ORACLE_SO, "orc")
    return _orc


def reference() -> _Lib:
    """The reference itself (oracle/_ref/libref_shim.so). Raises if it was never built."""
    global _ref
    if _ref is None:
        if not have_ref():
            raise RuntimeError("oracle/_ref/libref_shim.so missing: run `make -C oracle` where /root/reference exists")
        _ref = _Lib(REF_SHIM_SO, "ref")
    return _ref
