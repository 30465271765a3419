"""Parity of the CUDA POA / correction path (through the C ABI) against the UNMODIFIED reference
(oracle/_ref/libref_shim.so: spoa AVX2 engine + correct.cpp, compiled from /root/reference by oracle/Makefile; the
built .so travels to the GPU box) and against committed golden vectors made from it (tests/golden/make_golden_poa.py).
Bar: bit-exact — alignments (node,pos pairs), MSA rows, consensi.fq / uncorrected.fq bytes, corrected.fq as a multiset."""
import json
import os

import numpy as np
import pytest

from tools import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def pack(seed=3, n=12, length=400.0, **kw):
    return synth.generate(seed=seed, n_genes=1, reads_per_tx=n, len_mean=length, len_sd=0.0, len_min=int(length),
                          len_max=int(length), p_flip=0.0, shuffle=False, **kw)


@pytest.mark.parametrize("seed,n,length", [(1, 6, 120.0), (2, 16, 500.0), (3, 32, 2000.0), (4, 3, 60.0)])
def test_poa_msa_and_alignments_match_reference(ctx, ref, seed, n, length):
    rs = pack(seed, n, length)
    rows, alns = ctx.poa_msa(rs.bases, rs.offsets, want_alignments=True)
    erows, ealns = ref.poa_msa(rs.bases, rs.offsets, want_alignments=True)
    for i, (a, b) in enumerate(zip(alns, ealns)):
        assert np.array_equal(a, b), "alignment %d differs" % i
    assert rows == erows
    for i in range(rs.n):  # LocalAffineMSA property (spoa_test.cpp:495-516): row minus gaps == input
        assert rows[i].replace(b"-", b"") == rs.seq(i)
    st = ctx.stats()
    assert st["poa_alignments"] == rs.n - 1 and st["poa_cells"] > 0


def test_poa_high_error_and_divergent_reads(ctx, ref):
    """noisy reads + an unrelated read (max score tiny / branching graph, vertical and horizontal gap walks)"""
    rs = pack(7, 14, 600.0, p_sub=0.08, p_ins=0.06, p_del=0.06)
    rng = np.random.default_rng(0)
    seqs = [rs.seq(i) for i in range(rs.n)]
    seqs.insert(5, bytes(rng.choice(list(b"ACGT"), size=300).astype(np.uint8)))
    seqs.append(seqs[0][100:350])
    seqs.append(b"ACGT" * 40)
    rs2 = synth.from_sequences(seqs)
    rows, alns = ctx.poa_msa(rs2.bases, rs2.offsets, want_alignments=True)
    erows, ealns = ref.poa_msa(rs2.bases, rs2.offsets, want_alignments=True)
    for i, (a, b) in enumerate(zip(alns, ealns)):
        assert np.array_equal(a, b), "alignment %d differs" % i
    assert rows == erows


def test_poa_wide_path_long_read(ctx, ref):
    """reads long enough that scores leave int16 -> 32-bit kernel variant; several column chunks per row"""
    rs = pack(9, 4, 7000.0)
    rows, alns = ctx.poa_msa(rs.bases, rs.offsets, want_alignments=True)
    erows, ealns = ref.poa_msa(rs.bases, rs.offsets, want_alignments=True)
    for a, b in zip(alns, ealns):
        assert np.array_equal(a, b)
    assert rows == erows


def test_poa_golden_msa(ctx):
    path = os.path.join(GOLD, "poa_msa.json")
    if not os.path.exists(path):
        pytest.skip("golden not generated")
    for case in json.load(open(path)):
        rs = synth.from_sequences([s.encode() for s in case["seqs"]])
        rows = ctx.poa_msa(rs.bases, rs.offsets)
        assert [r.decode() for r in rows] == case["msa"]


def clusters_of(sizes):
    from rattle_b200 import ClusterSet
    off = np.zeros(len(sizes) + 1, np.int64)
    off[1:] = np.cumsum(sizes)
    n = int(off[-1])
    ids = np.arange(n, dtype=np.int32)
    return ClusterSet(ids[off[:-1]].copy(), np.zeros(len(sizes), np.uint8), off, ids, np.zeros(n, np.uint8))


def check_correct(out, exp):
    assert out[2] == exp[2], "consensi differ"
    assert out[1] == exp[1], "uncorrected differ"
    def recs(b):
        lines = b.split(b"\n")
        return sorted(tuple(lines[i:i + 4]) for i in range(0, len(lines) - 1, 4))
    assert recs(out[0]) == recs(exp[0]), "corrected differ"
    assert out[0] == exp[0], "corrected order differs from the reference's -t 1 order"


def test_correct_reads_matches_reference(ctx, ref):
    """config-4 shape, small: clusters x 12 forward reads x 500 nt + one tiny cluster (-> uncorrected)"""
    rs = synth.generate(seed=4, n_genes=6, reads_per_tx=12, len_mean=500.0, len_sd=0.0, len_min=500, len_max=500,
                        p_flip=0.0, shuffle=False)
    cl = clusters_of([12, 12, 12, 12, 12, 9, 3])
    out = ctx.correct_reads(rs.bases, rs.quals, rs.offsets, cl, min_reads=5)
    exp = ref.correct_reads(rs.bases, rs.quals, rs.offsets, cl.as_dict(), min_reads=5, n_threads=1)
    check_correct(out, exp)
    assert out[2].count(b"@gene_cluster_") == 6


def test_correct_reads_split_packs_rev_members_and_transcript_mode(ctx, ref):
    """split < cluster size -> several packs + third POA (correct.cpp:518-538); reverse members are
    reverse-complemented (correct.cpp:343-346); gene ids -> transcript_cluster headers"""
    from rattle_b200 import ClusterSet
    rs = synth.generate(seed=8, n_genes=2, reads_per_tx=30, len_mean=450.0, len_sd=0.0, len_min=450, len_max=450,
                        p_flip=0.5, shuffle=False)
    srs = rs.sorted_by_length()[0]
    clo = ctx.cluster_reads(srs.bases, srs.offsets, is_rna=False)
    assert clo.n_clusters == 2 and clo.mem_rev.any()
    cl = ClusterSet(clo.main_id, clo.main_rev, clo.cl_off, clo.mem_id, clo.mem_rev,
                    np.array([7, 9], np.int32), np.full(len(clo.mem_id), 3, np.int32))
    out = ctx.correct_reads(srs.bases, srs.quals, srs.offsets, cl, split=8, min_reads=2)
    exp = ref.correct_reads(srs.bases, srs.quals, srs.offsets, cl.as_dict(), gene_main=cl.main_gene, gene_mem=cl.mem_gene,
                            split=8, min_reads=2, n_threads=1)
    check_correct(out, exp)
    assert b"@transcript_cluster_0 gene_cluster_7" in out[2]


def test_correct_golden(ctx):
    path = os.path.join(GOLD, "correct_small.json")
    if not os.path.exists(path):
        pytest.skip("golden not generated")
    g = json.load(open(path))
    rs = synth.generate(**g["synth"])
    cl = clusters_of(g["sizes"])
    out = ctx.correct_reads(rs.bases, rs.quals, rs.offsets, cl, **g["kw"])
    import hashlib
    assert hashlib.sha256(out[2]).hexdigest() == g["consensi_sha256"]
    assert hashlib.sha256(out[1]).hexdigest() == g["uncorrected_sha256"]
    assert hashlib.sha256(out[0]).hexdigest() == g["corrected_sha256"]
    assert out[2].decode().splitlines()[:2] == g["consensi_head"]


# ---- int16 strip kernel (poa_strip_kernel.cuh): geometry and graph-shape edge cases, all against the reference
@pytest.mark.parametrize("length,n", [(5, 4), (9, 5), (255, 5), (256, 5), (257, 6), (2049, 5), (3500, 4)])
def test_poa_strip_boundaries(ctx, ref, length, n):
    """query lengths around the lane (8), strip (256) and pass (8 strips) boundaries; 3500 = two passes of 7 strips"""
    rs = pack(11 + length % 7, n, float(length), trunc_max=0)
    rows, alns = ctx.poa_msa(rs.bases, rs.offsets, want_alignments=True)
    erows, ealns = ref.poa_msa(rs.bases, rs.offsets, want_alignments=True)
    for i, (a, b) in enumerate(zip(alns, ealns)):
        assert np.array_equal(a, b), "alignment %d differs" % i
    assert rows == erows


def test_poa_strip_bushy_graph_spilled_rows(ctx, ref):
    """60 noisy reads: rows with more than 3 predecessors (overflow list) and predecessors further back than the
    shared-memory ring (spilled rows read back from HBM)"""
    rs = pack(21, 60, 700.0, p_sub=0.06, p_ins=0.04, p_del=0.04)
    rows, alns = ctx.poa_msa(rs.bases, rs.offsets, want_alignments=True)
    erows, ealns = ref.poa_msa(rs.bases, rs.offsets, want_alignments=True)
    for i, (a, b) in enumerate(zip(alns, ealns)):
        assert np.array_equal(a, b), "alignment %d differs" % i
    assert rows == erows


def test_poa_strip_mixed_lengths_and_letters(ctx, ref):
    """reads of very different lengths in one pack (several CTA widths over the steps), U instead of T, an unrelated
    read (no positive score -> empty alignment) and a read with N (not representable -> int32 kernel)"""
    rs = pack(31, 10, 900.0)
    seqs = [rs.seq(i) for i in range(rs.n)]
    seqs[3] = seqs[3][:300]
    seqs[4] = seqs[4][200:260]
    seqs[5] = seqs[5].replace(b"T", b"U")
    seqs.insert(6, b"G" * 40)
    rs2 = synth.from_sequences(seqs)
    rows, alns = ctx.poa_msa(rs2.bases, rs2.offsets, want_alignments=True)
    erows, ealns = ref.poa_msa(rs2.bases, rs2.offsets, want_alignments=True)
    for i, (a, b) in enumerate(zip(alns, ealns)):
        assert np.array_equal(a, b), "alignment %d differs" % i
    assert rows == erows
    seqs[2] = seqs[2][:100] + b"N" + seqs[2][101:]
    rs3 = synth.from_sequences(seqs)
    assert ctx.poa_msa(rs3.bases, rs3.offsets) == ref.poa_msa(rs3.bases, rs3.offsets)


def test_poa_int32_kernel_still_matches(ctx, ref):
    """option poa_kernel=1 forces the first-generation int32 kernel (the fallback for other scores / letters)"""
    rs = pack(5, 10, 600.0)
    ctx.set_option("poa_kernel", 1)
    try:
        rows, alns = ctx.poa_msa(rs.bases, rs.offsets, want_alignments=True)
    finally:
        ctx.set_option("poa_kernel", 0)
    erows, ealns = ref.poa_msa(rs.bases, rs.offsets, want_alignments=True)
    for a, b in zip(alns, ealns):
        assert np.array_equal(a, b)
    assert rows == erows


@pytest.mark.parametrize("opt,val", [("poa_gpu_sort", 0), ("poa_mirror_pct", 12), ("poa_mirror_pct", 30)])
def test_poa_host_sort_and_mirror_demotion(ctx, ref, opt, val):
    """graphs sorted on the host (poa_gpu_sort=0), and graphs that outgrow a deliberately small device mirror in the
    middle of the chain and are demoted to the host path (deferred sort caught up, records staged by the host)"""
    rs = pack(13, 24, 500.0, p_sub=0.05, p_ins=0.03, p_del=0.03)
    ctx.set_option(opt, val)
    try:
        rows, alns = ctx.poa_msa(rs.bases, rs.offsets, want_alignments=True)
        cl = clusters_of([24])
        out = ctx.correct_reads(rs.bases, rs.quals, rs.offsets, cl, min_reads=5)
    finally:
        ctx.set_option("poa_gpu_sort", 1)
        ctx.set_option("poa_mirror_pct", 100)
    erows, ealns = ref.poa_msa(rs.bases, rs.offsets, want_alignments=True)
    for i, (a, b) in enumerate(zip(alns, ealns)):
        assert np.array_equal(a, b), "alignment %d differs" % i
    assert rows == erows
    check_correct(out, ref.correct_reads(rs.bases, rs.quals, rs.offsets, cl.as_dict(), min_reads=5, n_threads=1))


# ---- MSA post-processing on the GPU (poa_vote.cuh) against the host pipeline and the reference
@pytest.mark.parametrize("quals", ["random", "constant"])
def test_correct_reads_device_vote_equals_host_vote_and_reference(ctx, ref, quals):
    """fix_msa_ends / column vote / read correction / consensus as kernels (option poa_device_vote, the default) give the
    bytes of the host pipeline (poa_device_vote=0) and of the unmodified reference; constant qualities make every mean
    error a table entry or a sum of equal terms — the columns whose quality symbol the device hands to the host's log10;
    ragged read lengths put short, badly aligned read ends into the MSAs (fix_msa_ends trims them)"""
    rs = synth.generate(seed=23, n_genes=5, reads_per_tx=14, len_mean=700.0, len_sd=250.0, len_min=200, len_max=1500,
                        p_flip=0.0, shuffle=False, p_sub=0.05, p_ins=0.04, p_del=0.04)
    if quals == "constant":
        rs.quals[:] = ord("5")
    # clusters = the genes (14 reads each) with reads in length-descending order, one cluster too small to correct
    lens = rs.lengths()
    order = np.concatenate([np.arange(g * 14, (g + 1) * 14)[np.argsort(-lens[g * 14:(g + 1) * 14], kind="stable")] for g in range(5)])
    rs = rs.take(order)
    cl = clusters_of([14, 14, 14, 14, 10, 4])
    exp = ref.correct_reads(rs.bases, rs.quals, rs.offsets, cl.as_dict(), min_reads=5, n_threads=1)
    outs = {}
    for vote in (1, 0):
        ctx.set_option("poa_device_vote", vote)
        try:
            outs[vote] = ctx.correct_reads(rs.bases, rs.quals, rs.offsets, cl, min_reads=5)
        finally:
            ctx.set_option("poa_device_vote", 1)
        check_correct(outs[vote], exp)
    assert outs[0] == outs[1]
