"""CPU: pins the clustering oracle (oracle/rattle_oracle.cpp) against golden vectors produced by the UNMODIFIED
reference (tests/golden/cluster_small.json, config2_*.json) and, where oracle/_ref was built, against the reference's
own functions on fresh seeded inputs."""
import json
import os

import numpy as np
import pytest

from tools import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gold():
    return json.load(open(os.path.join(GOLD, "cluster_small.json")))


def test_pair_match_golden(orc, gold):
    seqs = [s.encode() for s in gold["seqs"]]
    for case in gold["pairs"]:
        got = []
        for i in range(len(seqs)):
            for j in range(i + 1, len(seqs)):
                got.append(orc.pair_match(seqs[i], seqs[j], case["k"], case["t_s"], case["t_v"], case["thr"], case["is_rna"]))
        assert got == case["match"]


def test_extract_kmers_golden(orc, gold):
    seqs = [s.encode() for s in gold["seqs"]]
    for i, g in enumerate(gold["kmers_k10"]):
        n, fh, fp, rh, rp, bf, br = orc.extract_kmers(seqs[i], 10, True)
        assert n == len(seqs[i]) - 10  # the last k-mer is dropped (kmer.cpp:9-10)
        assert fh.tolist() == g["fh"] and fp.tolist() == g["fp"] and rh.tolist() == g["rh"] and rp.tolist() == g["rp"]
        assert [int(x) for x in bf] == g["bf"] and [int(x) for x in br] == g["br"]


def test_cluster_reads_golden(orc, gold):
    for case in gold["cluster"]:
        rs = synth.generate(**case["synth"]).sorted_by_length()[0]
        cl = orc.cluster_reads(rs.bases, rs.offsets, is_rna=case["is_rna"], n_threads=4, **case["kw"])
        assert cl["n_clusters"] == case["n_clusters"]
        for k in ("main_id", "main_rev", "cl_off", "mem_id", "mem_rev"):
            assert cl[k].tolist() == case[k], k


def test_cluster_reads_config2_20k_digest(orc):
    """20 k reads of the bench workload: the oracle reproduces the unmodified reference's digest"""
    import hashlib
    gold = json.load(open(os.path.join(GOLD, "config2_400.json")))
    rs = synth.config2(n_genes=400).sorted_by_length()[0]
    cl = orc.cluster_reads(rs.bases, rs.offsets, is_rna=False, n_threads=os.cpu_count() or 1)
    h = hashlib.sha256()
    for k in ("main_id", "main_rev", "cl_off", "mem_id", "mem_rev"):
        h.update(np.ascontiguousarray(cl[k]).tobytes())
    assert cl["n_clusters"] == gold["n_clusters"] and h.hexdigest() == gold["sha256"]


def test_var_edge_cases(orc):
    assert orc.var([]) == 0.0  # utils.cpp:41
    assert np.isnan(orc.var([5]))  # n=1 -> 0/0 (utils.cpp:54): the accept test `var < t_v` then rejects
    assert orc.var([1, 1, 1]) == 0.0
    assert orc.var([1, 2, 3, 4]) == pytest.approx(5.0 / 3.0)


def test_similarity_known_answers(orc):
    # empty -> 0 bases; one match -> k bases, no distances
    assert orc.similarity([], [], 10)[0] == 0
    b, d = orc.similarity([5], [9], 10)
    assert b == 10 and len(d) == 0
    # co-linear run of adjacent k-mers: each adds one new base (overlap k-1)
    b, d = orc.similarity([0, 1, 2, 3], [7, 8, 9, 10], 10)
    assert b == 13 and d.tolist() == [0, 0, 0]
    # a far-away hit on the second read only (gap in one coordinate < k, other >= k) is dropped by the chain filter
    b, d = orc.similarity([0, 1, 2], [0, 1, 50], 10)
    assert b == 11 and d.tolist() == [0]


def test_hps_encode_known_bytes(orc):
    """varint count, zig-zag ids, rev byte, zig-zag gene (-1 -> 0x01)"""
    cl = dict(n_clusters=1, main_id=np.array([1640], np.int32), main_rev=np.array([0], np.uint8),
              cl_off=np.array([0, 2], np.int64), mem_id=np.array([1640, 1151], np.int32), mem_rev=np.array([0, 1], np.uint8))
    assert orc.hps_encode(cl).hex() == "01" + "d01900" + "01" + "02" + "d01900" + "01" + "fe1101" + "01"


# ---- against the reference itself (only where oracle/_ref was built)
def test_functions_match_reference(orc, ref):
    rs = synth.generate(seed=17, n_genes=5, reads_per_tx=6, len_mean=500.0, len_sd=150.0, len_min=100, len_max=1200)
    for k, both in [(10, True), (6, False), (16, True)]:
        for i in range(0, rs.n, 5):
            a = orc.extract_kmers(rs.seq(i), k, both)
            b = ref.extract_kmers(rs.seq(i), k, both)
            assert a[0] == b[0]
            for x, y in zip(a[1:], b[1:]):
                assert np.array_equal(x, y)
    km = [orc.extract_kmers(rs.seq(i), 10, True) for i in range(rs.n)]
    rng = np.random.default_rng(3)
    for _ in range(40):
        i, j = rng.integers(0, rs.n, 2)
        h2, p2 = (km[j][3], km[j][4]) if rng.random() < 0.5 else (km[j][1], km[j][2])
        fa, sa = orc.common_kmers(km[i][1], km[i][2], h2, p2)
        fb, sb = ref.common_kmers(km[i][1], km[i][2], h2, p2)
        assert np.array_equal(fa, fb) and np.array_equal(sa, sb)
        ba, da = orc.similarity(fa, sa, 10)
        bb, db = ref.similarity(fb, sb, 10)
        assert ba == bb and np.array_equal(da, db)
        va, vb = orc.var(da), ref.var(db)
        assert va == vb or (np.isnan(va) and np.isnan(vb))


def test_similarity_random_matches_reference(orc, ref):
    rng = np.random.default_rng(11)
    for n in (1, 2, 3, 17, 200, 1500):
        first = np.sort(rng.integers(0, 2000, n)).astype(np.int32)
        second = rng.integers(0, 2000, n).astype(np.int32)
        order = np.lexsort((second, first))
        first, second = first[order], second[order]
        for k in (6, 10, 11):
            ba, da = orc.similarity(first, second, k)
            bb, db = ref.similarity(first, second, k)
            assert ba == bb and np.array_equal(da, db)


@pytest.mark.parametrize("is_rna", [False, True])
def test_cluster_reads_matches_reference(orc, ref, is_rna):
    rs = synth.generate(seed=7, n_genes=25, reads_per_tx=12).sorted_by_length()[0]
    a = orc.cluster_reads(rs.bases, rs.offsets, is_rna=is_rna, n_threads=8)
    b = ref.cluster_reads(rs.bases, rs.offsets, is_rna=is_rna, n_threads=8)
    assert a["n_clusters"] == b["n_clusters"]
    for k in ("main_id", "main_rev", "cl_off", "mem_id", "mem_rev"):
        assert np.array_equal(a[k], b[k])


def test_toyset_pin(orc):
    """`cluster --rna` on toyset/rna gives 546 clusters (config 1). Needs /root/reference (skipped on the GPU box)."""
    fq = "/root/reference/toyset/rna/input/sample.fastq"
    if not os.path.exists(fq):
        pytest.skip("toyset not available")
    rs = synth.read_fastq(fq)
    keep = [i for i in range(rs.n) if 150 <= len(rs.seq(i)) <= 100000 and b"N" not in rs.seq(i)]  # fasta.cpp:301-340
    rs = rs.take(keep).sorted_by_length()[0]
    cl = orc.cluster_reads(rs.bases, rs.offsets, is_rna=True, n_threads=os.cpu_count() or 1)
    assert rs.n == 8304 and cl["n_clusters"] == 546
