"""CPU: the C-ABI library loads and exports every symbol include/rattle_b200.h declares (no compute calls without a
GPU), the product never touches oracle/, the clusters.out codec round-trips and matches the oracle's encoder."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "rattle_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rtl_[a-z_0-9]+)\s*\(", src)) - {"rtl_allreduce_min_fn"})


def test_library_exports_every_declared_symbol():
    import rattle_b200
    from rattle_b200.api import SYMBOLS
    lib = rattle_b200.load_library()
    decl = header_symbols()
    assert decl == sorted(SYMBOLS)
    out = subprocess.run(["nm", "-D", "--defined-only", rattle_b200.lib_path()], capture_output=True, text=True).stdout
    exported = set(l.split()[-1] for l in out.splitlines() if l.strip())
    for s in decl:
        assert s in exported, s
        assert hasattr(lib, s)


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import rattle_b200
    with pytest.raises(rattle_b200.RattleError) as e:
        rattle_b200.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_product_never_uses_the_oracle():
    bad = []
    for top in ("rattle_b200", "integration", "include"):  # the library, the drop-in CLI shim, the ABI
        for dp, _, files in os.walk(os.path.join(ROOT, top)):
            if "build" in dp.split(os.sep) or "_build" in dp.split(os.sep):
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) or f == "Makefile":
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"import\s+oracle|from\s+oracle|liboracle|libref_shim|rattle_oracle\.h|orc_[a-z_]+\(|oracle/", txt):
                        bad.append(os.path.join(top, f))
    assert not bad, bad
    # the drop-in CLI links the library and the reference's main/fasta/utils objects only: no spoa, no reference kernels
    dropin = os.path.join(ROOT, "integration", "_build", "rattle")
    if os.path.exists(dropin):
        out = subprocess.run(["nm", "-C", "--defined-only", dropin], capture_output=True, text=True).stdout
        assert "spoa::" not in out and "extract_kmers_from_read" not in out and "cluster_together" not in out
        assert "rattle_b200" in subprocess.run(["ldd", dropin], capture_output=True, text=True).stdout
    # and the shared library has no dependency on them
    import rattle_b200
    out = subprocess.run(["ldd", rattle_b200.lib_path()], capture_output=True, text=True).stdout
    assert "oracle" not in out and "ref_shim" not in out


def test_hps_codec_roundtrip_and_matches_oracle(orc):
    import rattle_b200
    rng = np.random.default_rng(0)
    sizes = rng.integers(1, 40, 50)
    off = np.zeros(51, np.int64)
    off[1:] = np.cumsum(sizes)
    n = int(off[-1])
    cl = rattle_b200.ClusterSet(rng.integers(0, 10 ** 6, 50).astype(np.int32), rng.integers(0, 2, 50).astype(np.uint8), off,
                                rng.integers(0, 2 ** 31 - 1, n).astype(np.int32), rng.integers(0, 2, n).astype(np.uint8),
                                rng.integers(-1, 300, 50).astype(np.int32), rng.integers(-1, 300, n).astype(np.int32))
    buf = rattle_b200.hps_encode(cl)
    assert buf == orc.hps_encode(cl.as_dict(), cl.main_gene, cl.mem_gene)
    back = rattle_b200.hps_decode(buf)
    for k in ("main_id", "main_rev", "cl_off", "mem_id", "mem_rev", "main_gene", "mem_gene"):
        assert np.array_equal(getattr(back, k), getattr(cl, k)), k
    with pytest.raises(rattle_b200.RattleError):
        rattle_b200.hps_decode(buf[:len(buf) // 2])


def test_hps_known_bytes():
    """clusters.out prefix measured on the reference's toyset run: a2 04 | d0 19 00 01 | 05 | ..."""
    import rattle_b200
    cl = rattle_b200.ClusterSet(np.array([1640], np.int32), np.array([0], np.uint8), np.array([0, 2], np.int64),
                                np.array([1640, 1151], np.int32), np.array([0, 0], np.uint8))
    assert rattle_b200.hps_encode(cl).hex().startswith("01d019000102d0190001fe110001")
    head = bytes.fromhex("a204")  # varint 546
    assert int(head[0] & 0x7f) | (head[1] << 7) == 546


def test_correct_reads_wrapper_buffers_with_a_mock_library():
    """rattle_b200.Context.correct_reads without a GPU: the ctypes plumbing (buffer reuse, capacity retry, views vs
    bytes) against a mock of rtl_correct_reads that fills the three caller buffers."""
    import ctypes
    import rattle_b200
    from rattle_b200.api import Context

    class Mock:
        def __init__(self):
            self.calls = 0
            self.payload = [b"@r0\\nACGT\\n+\\nIIII\\n" * 3, b"", b"@gene_cluster_0 reads=3 labels=\\nACGT\\n+\\nKKKK\\n"]

        def rtl_correct_reads(self, *args):
            self.calls += 1
            outs = args[-6:]
            rc = 0
            for i in range(3):
                buf, ln = outs[2 * i], outs[2 * i + 1]._obj
                need = len(self.payload[i])
                if need > ln.value:
                    rc = -3
                else:
                    ctypes.memmove(buf, self.payload[i], need)
                ln.value = need
            return rc

        def rtl_set_cluster_ids(self, h, ids, n):
            return 0

        def rtl_last_error(self, h):
            return b"mock"

    ctx = object.__new__(Context)
    ctx.L, ctx.h = Mock(), None
    cl = rattle_b200.ClusterSet(np.array([0], np.int32), np.array([0], np.uint8), np.array([0, 3], np.int64),
                                np.arange(3, dtype=np.int32), np.zeros(3, np.uint8))
    bases = np.frombuffer(b"ACGT" * 3, np.uint8)
    offs = np.array([0, 4, 8, 12], np.uint64)
    out = ctx.correct_reads(bases, bases, offs, cl)
    assert out == tuple(ctx.L.payload) and all(isinstance(x, bytes) for x in out)
    first = ctx._corr_bufs
    views = ctx.correct_reads(bases, bases, offs, cl, as_bytes=False)
    assert ctx._corr_bufs is first  # buffers are reused
    assert [v.tobytes() for v in views] == ctx.L.payload
    ctx.L.payload[0] = b"x" * (len(first[0]) + 100)  # larger than the held buffers: one retry with the reported size
    n = ctx.L.calls
    views = ctx.correct_reads(bases, bases, offs, cl, as_bytes=False)
    assert ctx.L.calls == n + 2 and len(views[0]) == len(ctx.L.payload[0])


@pytest.mark.parametrize("name,n_clusters,gene_level", [("clusters_rna_1500.out", None, True),
                                                       ("clusters_rna_iso_1500.out", None, False)])
def test_hps_codec_on_reference_written_clusters_out(name, n_clusters, gene_level):
    """clusters.out files WRITTEN BY THE REFERENCE CLI (oracle/_ref/rattle cluster [--iso] on the 1500-read fixture;
    tests/golden/make_golden_big.py hps): decode -> encode reproduces the file byte for byte, and the decoded set is a
    partition of read indices with gene ids as main.cpp:266-273 / :302-318 write them."""
    import rattle_b200
    path = os.path.join(ROOT, "tests", "golden", name)
    buf = open(path, "rb").read()
    cs = rattle_b200.hps_decode(buf)
    assert rattle_b200.hps_encode(cs) == buf
    assert cs.n_clusters > 100 and cs.cl_off[0] == 0 and cs.cl_off[-1] == len(cs.mem_id)
    assert len(set(cs.mem_id.tolist())) == len(cs.mem_id)  # every read in exactly one cluster
    for c in range(cs.n_clusters):
        assert cs.main_id[c] in cs.mem_id[cs.cl_off[c]:cs.cl_off[c + 1]]
    if gene_level:
        assert (cs.main_gene == -1).all() and (cs.mem_gene == -1).all()
    else:  # --iso: transcript clusters carry the id of the gene cluster they were split from
        assert (cs.main_gene >= 0).all() and cs.main_gene.max() < cs.n_clusters
        assert (np.diff(cs.main_gene) >= 0).all()
