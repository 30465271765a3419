"""CPU: pins the POA/correction oracle (oracle/poa_oracle.cpp) against golden vectors from the UNMODIFIED reference
(tests/golden/poa_msa.json incl. 12 reads of spoa's own test data, correct_small.json) and against the reference live."""
import hashlib
import json
import os

import numpy as np
import pytest

from tools import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_poa_msa_golden(orc):
    for case in json.load(open(os.path.join(GOLD, "poa_msa.json"))):
        rs = synth.from_sequences([s.encode() for s in case["seqs"]])
        rows = orc.poa_msa(rs.bases, rs.offsets)
        assert [r.decode() for r in rows] == case["msa"]
        for r, s in zip(rows, case["seqs"]):  # spoa_test.cpp:495-516 LocalAffineMSA property
            assert r.replace(b"-", b"").decode() == s


def test_correct_golden(orc):
    g = json.load(open(os.path.join(GOLD, "correct_small.json")))
    rs = synth.generate(**g["synth"])
    sizes = g["sizes"]
    off = np.zeros(len(sizes) + 1, np.int64)
    off[1:] = np.cumsum(sizes)
    ids = np.arange(int(off[-1]), dtype=np.int32)
    cl = dict(n_clusters=len(sizes), main_id=ids[off[:-1]].copy(), main_rev=np.zeros(len(sizes), np.uint8), cl_off=off,
              mem_id=ids, mem_rev=np.zeros(len(ids), np.uint8))
    out = orc.correct_reads(rs.bases, rs.quals, rs.offsets, cl, **g["kw"])
    assert hashlib.sha256(out[2]).hexdigest() == g["consensi_sha256"]
    assert hashlib.sha256(out[1]).hexdigest() == g["uncorrected_sha256"]
    assert hashlib.sha256(out[0]).hexdigest() == g["corrected_sha256"]


def test_poa_matches_reference_live(orc, ref):
    for seed, n, L in [(5, 7, 200.0), (6, 20, 450.0)]:
        rs = synth.generate(seed=seed, n_genes=1, reads_per_tx=n, len_mean=L, len_sd=0.0, len_min=int(L), len_max=int(L),
                            p_flip=0.0, shuffle=False, p_sub=0.06, p_ins=0.04, p_del=0.04)
        a, aa = orc.poa_msa(rs.bases, rs.offsets, want_alignments=True)
        b, bb = ref.poa_msa(rs.bases, rs.offsets, want_alignments=True)
        assert a == b
        for x, y in zip(aa, bb):
            assert np.array_equal(x, y)


def test_correct_split_and_rev_matches_reference_live(orc, ref):
    rs = synth.generate(seed=8, n_genes=2, reads_per_tx=14, len_mean=300.0, len_sd=0.0, len_min=300, len_max=300,
                        p_flip=0.5, shuffle=False)
    n = rs.n
    cl = dict(n_clusters=2, main_id=np.array([0, 14], np.int32), main_rev=np.zeros(2, np.uint8),
              cl_off=np.array([0, 14, 28], np.int64), mem_id=np.arange(n, dtype=np.int32), mem_rev=rs.truth_rev.copy())
    gm = np.array([4, 6], np.int32)
    gs = np.full(n, 1, np.int32)
    a = orc.correct_reads(rs.bases, rs.quals, rs.offsets, cl, gene_main=gm, gene_mem=gs, split=5, min_reads=2)
    b = ref.correct_reads(rs.bases, rs.quals, rs.offsets, cl, gene_main=gm, gene_mem=gs, split=5, min_reads=2, n_threads=1)
    assert a == b
    assert b"@transcript_cluster_1 gene_cluster_6" in a[2]
