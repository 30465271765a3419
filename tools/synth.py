"""ctypes front end of tools/synth.c plus small read-set helpers shared by tests/ and bench.py.

A ReadSet is the flat layout every C-ABI entry point takes: `bases` (uint8 ASCII, concatenated), `offsets`
(uint64, n+1), optional `quals` (uint8, same shape as bases).
"""
import ctypes
import os
import subprocess
from dataclasses import dataclass
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libsynth.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "synth.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", _LIB, src, "-lm"])
    return _LIB


_lib = None


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_LIB)
        lib.synth_create.restype = ctypes.c_void_p
        lib.synth_create.argtypes = [ctypes.c_uint64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                     ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_int]
        lib.synth_n_reads.restype = ctypes.c_uint32
        lib.synth_n_reads.argtypes = [ctypes.c_void_p]
        lib.synth_total_bases.restype = ctypes.c_uint64
        lib.synth_total_bases.argtypes = [ctypes.c_void_p]
        lib.synth_copy.restype = None
        lib.synth_copy.argtypes = [ctypes.c_void_p] + [ctypes.c_void_p] * 5
        lib.synth_free.restype = None
        lib.synth_free.argtypes = [ctypes.c_void_p]
        lib.synth_take.restype = None
        lib.synth_take.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p,
                                   ctypes.c_void_p]
        _lib = lib
    return _lib


@dataclass
class ReadSet:
    bases: np.ndarray  # uint8
    offsets: np.ndarray  # uint64, n+1
    quals: Optional[np.ndarray] = None
    truth_tx: Optional[np.ndarray] = None
    truth_rev: Optional[np.ndarray] = None

    @property
    def n(self) -> int:
        return len(self.offsets) - 1

    def lengths(self) -> np.ndarray:
        return np.diff(self.offsets.astype(np.int64))

    def seq(self, i: int) -> bytes:
        return self.bases[int(self.offsets[i]):int(self.offsets[i + 1])].tobytes()

    def qual(self, i: int) -> bytes:
        return self.quals[int(self.offsets[i]):int(self.offsets[i + 1])].tobytes()

    def take(self, idx) -> "ReadSet":
        """Read set made of reads idx (in that order)."""
        idx = np.asarray(idx, dtype=np.int64)
        lens = self.lengths()[idx]
        offs = np.zeros(len(idx) + 1, dtype=np.uint64)
        offs[1:] = np.cumsum(lens)
        # one memcpy per read (tools/synth.c): an index per base would take 12 GB for a million 1.5-kb reads
        lib = _load()
        idx = np.ascontiguousarray(idx)
        src_off = np.ascontiguousarray(self.offsets, dtype=np.uint64)

        def gather(a):
            a = np.ascontiguousarray(a)
            out = np.empty(int(offs[-1]), dtype=a.dtype)
            lib.synth_take(a.ctypes.data, src_off.ctypes.data, idx.ctypes.data, len(idx), offs.ctypes.data, out.ctypes.data)
            return out
        return ReadSet(gather(self.bases), offs, None if self.quals is None else gather(self.quals),
                       None if self.truth_tx is None else self.truth_tx[idx],
                       None if self.truth_rev is None else self.truth_rev[idx])

    def sorted_by_length(self):
        """Stable length-descending order (fasta.cpp:458-464). Returns (sorted set, permutation)."""
        perm = np.argsort(-self.lengths(), kind="stable")
        return self.take(perm), perm


def generate(seed=42, n_genes=2000, n_isoforms=1, reads_per_tx=50, len_mean=1500.0, len_sd=150.0, len_min=600,
             len_max=3000, exon_skip=120, trunc_max=50, p_sub=0.03, p_ins=0.02, p_del=0.02, p_flip=0.5,
             shuffle=True) -> ReadSet:
    lib = _load()
    h = lib.synth_create(seed, n_genes, n_isoforms, reads_per_tx, len_mean, len_sd, len_min, len_max, exon_skip,
                         trunc_max, p_sub, p_ins, p_del, p_flip, 1 if shuffle else 0)
    try:
        n = lib.synth_n_reads(h)
        total = lib.synth_total_bases(h)
        bases = np.empty(total, dtype=np.uint8)
        quals = np.empty(total, dtype=np.uint8)
        offsets = np.empty(n + 1, dtype=np.uint64)
        tx = np.empty(n, dtype=np.int32)
        rev = np.empty(n, dtype=np.uint8)
        lib.synth_copy(h, bases.ctypes.data, quals.ctypes.data, offsets.ctypes.data, tx.ctypes.data, rev.ctypes.data)
    finally:
        lib.synth_free(h)
    return ReadSet(bases, offsets, quals, tx, rev)


# named workload shapes (SURVEY.md §8d / BASELINE.json configs)
def config2(n_genes=2000, seed=42) -> ReadSet:
    """cDNA gene clustering: n_genes transcripts x 50 reads x ~1.5 kb, both strands (2000 -> 100 k reads)."""
    return generate(seed=seed, n_genes=n_genes, n_isoforms=1, reads_per_tx=50)


def config3(n_genes=10000, seed=42) -> ReadSet:
    """cDNA, genes x 2 isoforms (120-nt exon skip) x 50 reads (10000 -> 1 M reads)."""
    return generate(seed=seed, n_genes=n_genes, n_isoforms=2, reads_per_tx=50)


def config4(n_clusters=10000, reads_per=32, seed=42) -> ReadSet:
    """correct: clusters x 32 forward-strand reads x 2 kb, emitted cluster by cluster (no shuffle)."""
    return generate(seed=seed, n_genes=n_clusters, n_isoforms=1, reads_per_tx=reads_per, len_mean=2000.0, len_sd=0.0,
                    len_min=2000, len_max=2000, p_flip=0.0, shuffle=False)


def config5(n_genes=10000, seed=42) -> ReadSet:
    """direct RNA: forward strand only, 4/3/3 % sub/ins/del."""
    return generate(seed=seed, n_genes=n_genes, n_isoforms=2, reads_per_tx=50, p_sub=0.04, p_ins=0.03, p_del=0.03,
                    p_flip=0.0)


def write_fastq(rs: ReadSet, path: str, prefix="r"):
    with open(path, "wb") as f:
        for i in range(rs.n):
            s = rs.seq(i)
            q = rs.qual(i) if rs.quals is not None else b"I" * len(s)
            f.write(b"@" + prefix.encode() + str(i).encode() + b"\n" + s + b"\n+\n" + q + b"\n")


def read_fastq(path: str) -> ReadSet:
    seqs, quals = [], []
    with open(path, "rb") as f:
        while True:
            h = f.readline()
            if not h:
                break
            seqs.append(f.readline().rstrip(b"\r\n"))
            f.readline()
            quals.append(f.readline().rstrip(b"\r\n"))
    return from_sequences(seqs, quals)


def from_sequences(seqs, quals=None) -> ReadSet:
    offs = np.zeros(len(seqs) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(s) for s in seqs])
    bases = np.frombuffer(b"".join(seqs), dtype=np.uint8).copy()
    q = None if quals is None else np.frombuffer(b"".join(quals), dtype=np.uint8).copy()
    return ReadSet(bases, offs, q)
