"""Parity of the CUDA clustering path (through the C ABI) against the CPU oracle on the same seeded inputs.
Bar: bit-exact (k-mer lists, bitvectors, popcounts, pass flags, bases, distances count, accept flags, cluster sets);
the variance double is compared with == as well (same IEEE operations in the same order)."""
import numpy as np
import pytest

from tools import synth

pytestmark = pytest.mark.gpu


def small_set(seed=3, genes=12, per=10, **kw):
    rs = synth.generate(seed=seed, n_genes=genes, reads_per_tx=per, **kw)
    return rs.sorted_by_length()[0]


def oracle_kmers(orc, rs, k, both):
    out = []
    for i in range(rs.n):
        out.append(orc.extract_kmers(rs.seq(i), k, both))
    return out


@pytest.mark.parametrize("k,both", [(10, True), (11, True), (6, False), (16, True), (3, True)])
def test_extract_kmers_matches_oracle(ctx, orc, k, both):
    rs = small_set(len_mean=700.0, len_sd=300.0, len_min=60, len_max=2500)
    fh, fp, rh, rp, bf, br = ctx.extract_kmers(rs.bases, rs.offsets, k, both)
    for i in range(rs.n):
        n, ofh, ofp, orh, orp, obf, obr = orc.extract_kmers(rs.seq(i), k, both)
        o = int(rs.offsets[i]) - i * k
        assert n == len(rs.seq(i)) - k
        assert np.array_equal(fh[o:o + n], ofh) and np.array_equal(fp[o:o + n], ofp), i
        assert np.array_equal(bf[i], obf), i
        if both:
            assert np.array_equal(rh[o:o + n], orh) and np.array_equal(rp[o:o + n], orp), i
            assert np.array_equal(br[i], obr), i
        else:
            assert not br[i].any()


def test_extract_kmers_long_reads(ctx, orc):
    """reads whose lists exceed every shared-memory sort class (global-memory bitonic path) and U bases"""
    rng = np.random.default_rng(5)
    seqs = [bytes(rng.choice(list(b"ACGU"), size=n).astype(np.uint8)) for n in (17000, 40000, 20, 1025 + 10, 16384 + 10)]
    rs = synth.from_sequences(seqs)
    k = 10
    fh, fp, rh, rp, bf, br = ctx.extract_kmers(rs.bases, rs.offsets, k, True)
    for i in range(rs.n):
        n, ofh, ofp, orh, orp, obf, obr = orc.extract_kmers(rs.seq(i), k, True)
        o = int(rs.offsets[i]) - i * k
        assert np.array_equal(fh[o:o + n], ofh) and np.array_equal(fp[o:o + n], ofp)
        assert np.array_equal(rh[o:o + n], orh) and np.array_equal(rp[o:o + n], orp)
        assert np.array_equal(bf[i], obf) and np.array_equal(br[i], obr)


def test_extract_rejects_bad_input(ctx):
    import rattle_b200
    rs = synth.from_sequences([b"ACGTACGTACGTNACGT", b"ACGTACGTACGTACGTT"])
    with pytest.raises(rattle_b200.RattleError):
        ctx.extract_kmers(rs.bases, rs.offsets, 10, True)
    rs = synth.from_sequences([b"ACGTACGTAC", b"ACGTACGTACGTACGTT"])  # len == k
    with pytest.raises(rattle_b200.RattleError):
        ctx.extract_kmers(rs.bases, rs.offsets, 10, True)


@pytest.mark.parametrize("thr", [0.4, 0.30000000000000004, 0.20000000000000007, 0.0])
def test_bv_scan_matches_numpy(ctx, orc, thr):
    rs = small_set(seed=9, genes=20, per=12)
    k = 10
    ctx.upload(rs.bases, rs.offsets)
    seeds = np.arange(0, rs.n, 3, dtype=np.int32)[:150]  # > one 128-seed tile
    targets = np.arange(rs.n, dtype=np.int32)
    cf, cr, passed = ctx.bv_scan(seeds, targets, thr, kmer_size=k, is_rna=False)
    bvs = [orc.extract_kmers(rs.seq(i), k, True) for i in range(rs.n)]
    bf = np.stack([b[5] for b in bvs]); br = np.stack([b[6] for b in bvs])
    pc = np.array([sum(bin(int(w)).count("1") for w in b) for b in bf])
    popc = np.vectorize(lambda w: bin(int(w)).count("1"))
    for si, s in enumerate(seeds):
        ecf = popc(bf[s][None, :] & bf).sum(1)
        ecr = popc(bf[s][None, :] & br).sum(1)
        assert np.array_equal(cf[si], ecf) and np.array_equal(cr[si], ecr)
        mmax = np.maximum(pc[s], pc).astype(np.float64)
        ef = (thr == 0) | (ecf.astype(np.float64) / mmax >= thr)
        er = ecr.astype(np.float64) / mmax >= thr
        assert np.array_equal(passed[si] & 1, ef.astype(np.uint8))
        assert np.array_equal(passed[si] >> 1, er.astype(np.uint8))


@pytest.mark.parametrize("k,t_s,t_v", [(10, 0.2, 1e6), (11, 0.3, 25.0), (6, 0.5, 25.0)])
def test_pair_similarity_matches_oracle(ctx, orc, k, t_s, t_v):
    rs = small_set(seed=11, genes=6, per=8, len_mean=900.0)
    ctx.upload(rs.bases, rs.offsets)
    n = rs.n
    a, b, st = [], [], []
    for i in range(n):
        for j in range(i + 1, n, 3):
            for s in (0, 1):
                a.append(i); b.append(j); st.append(s)
    res = ctx.pair_similarity(a, b, st, kmer_size=k, is_rna=False, t_s=t_s, t_v=t_v)
    km = [orc.extract_kmers(rs.seq(i), k, True) for i in range(n)]
    lens = rs.lengths()
    n_heavy = 0
    for t in range(len(a)):
        i, j, s = a[t], b[t], st[t]
        h1, p1 = km[i][1], km[i][2]
        h2, p2 = (km[j][3], km[j][4]) if s else (km[j][1], km[j][2])
        first, second = orc.common_kmers(h1, p1, h2, p2)
        assert res["n_common"][t] == len(first)
        bases, dist = orc.similarity(first, second, k)
        var = orc.var(dist)
        mn = float(min(lens[i], lens[j]))
        exp_acc = (float(bases) / mn >= t_s) and (var < t_v)
        if res["bases"][t] < 0:  # rejected by the exact bound, the reference rejects as well
            assert not exp_acc
            assert float(k * len(first)) / mn < t_s
        else:
            n_heavy += 1
            assert res["bases"][t] == bases and res["n_dist"][t] == len(dist)
            assert (res["var"][t] == var) or (np.isnan(res["var"][t]) and np.isnan(var))
        assert bool(res["accept"][t]) == exp_acc
    assert n_heavy > 20


def test_pair_similarity_repeats_and_scratch(ctx, orc):
    """low-complexity reads: quadratic cross pairs (kmer.cpp:56-61), matches spill to the global scratch"""
    rng = np.random.default_rng(2)
    core = bytes(rng.choice(list(b"ACGT"), size=300).astype(np.uint8))
    seqs = [core + b"A" * 120 + core[:50], b"A" * 90 + core, core[::-1] + b"AC" * 80, b"AC" * 100 + core]
    rs = synth.from_sequences(seqs)
    ctx.upload(rs.bases, rs.offsets)
    a, b, st = [], [], []
    for i in range(4):
        for j in range(4):
            if i != j:
                for s in (0, 1):
                    a.append(i); b.append(j); st.append(s)
    k = 6
    res = ctx.pair_similarity(a, b, st, kmer_size=k, is_rna=False, t_s=0.0, t_v=1e9)
    km = [orc.extract_kmers(rs.seq(i), k, True) for i in range(4)]
    for t in range(len(a)):
        i, j, s = a[t], b[t], st[t]
        h2, p2 = (km[j][3], km[j][4]) if s else (km[j][1], km[j][2])
        first, second = orc.common_kmers(km[i][1], km[i][2], h2, p2)
        bases, dist = orc.similarity(first, second, k)
        assert res["n_common"][t] == len(first)
        assert res["bases"][t] == bases and res["n_dist"][t] == len(dist)
        var = orc.var(dist)
        assert (res["var"][t] == var) or (np.isnan(res["var"][t]) and np.isnan(var))
    assert res["n_common"].max() > 1024  # exercised the global-scratch path


def assert_same_clusters(a, b):
    b = b if isinstance(b, dict) else b.as_dict()
    a = a if isinstance(a, dict) else a.as_dict()
    assert a["n_clusters"] == b["n_clusters"]
    for key in ("main_id", "main_rev", "cl_off", "mem_id", "mem_rev"):
        assert np.array_equal(a[key], b[key]), key


@pytest.mark.parametrize("is_rna", [False, True])
@pytest.mark.parametrize("wave", [512, 7])
def test_cluster_reads_matches_oracle(ctx, orc, is_rna, wave):
    rs = synth.generate(seed=7, n_genes=40, reads_per_tx=25).sorted_by_length()[0]
    ctx.set_option("wave", wave)
    try:
        got = ctx.cluster_reads(rs.bases, rs.offsets, is_rna=is_rna)
    finally:
        ctx.set_option("wave", 512)
    exp = orc.cluster_reads(rs.bases, rs.offsets, is_rna=is_rna, n_threads=8)
    assert_same_clusters(got, exp)
    st = ctx.stats()
    assert st["bv_pairs"] > 0 and st["full_pairs"] > 0 and st["kernel_launches"] > 0


@pytest.mark.parametrize("params", [dict(kmer_size=11, t_s=0.3, t_v=25.0),  # --iso (main.cpp:300)
                                    dict(kmer_size=6, t_s=0.5, t_v=25.0, bv_threshold=0.4, min_bv_threshold=0.4)])  # polish
def test_cluster_reads_iso_and_polish_params(ctx, orc, params):
    rs = synth.generate(seed=13, n_genes=15, n_isoforms=2, reads_per_tx=12, len_mean=1000.0).sorted_by_length()[0]
    kw = dict(kmer_size=10, t_s=0.2, t_v=1e6, bv_threshold=0.4, min_bv_threshold=0.2, bv_falloff=0.05)
    kw.update(params)
    got = ctx.cluster_reads(rs.bases, rs.offsets, is_rna=False, **kw)
    exp = orc.cluster_reads(rs.bases, rs.offsets, k=kw["kmer_size"], t_s=kw["t_s"], t_v=kw["t_v"],
                            bv_thr=kw["bv_threshold"], bv_min=kw["min_bv_threshold"], bv_falloff=kw["bv_falloff"],
                            is_rna=False, n_threads=8)
    assert_same_clusters(got, exp)


def test_cluster_reads_small_task_buffer_chunks(ctx, orc):
    """a tiny candidate buffer forces phase B to run in many target chunks"""
    rs = synth.generate(seed=21, n_genes=30, reads_per_tx=20, len_mean=2300.0, len_sd=50.0, len_max=3000).sorted_by_length()[0]
    ctx.set_option("task_cap", 4096 * 16)
    ctx.set_option("wave", 32)
    try:
        got = ctx.cluster_reads(rs.bases, rs.offsets, is_rna=False)
    finally:
        ctx.set_option("task_cap", 32 << 20)
        ctx.set_option("wave", 512)
    exp = orc.cluster_reads(rs.bases, rs.offsets, is_rna=False, n_threads=8)
    assert_same_clusters(got, exp)


def test_cluster_single_read_and_singletons(ctx, orc):
    rng = np.random.default_rng(1)
    seqs = [bytes(rng.choice(list(b"ACGT"), size=n).astype(np.uint8)) for n in (900, 800, 700, 650)]
    rs = synth.from_sequences(seqs)
    got = ctx.cluster_reads(rs.bases, rs.offsets)
    exp = orc.cluster_reads(rs.bases, rs.offsets)
    assert_same_clusters(got, exp)
    one = synth.from_sequences(seqs[:1])
    got = ctx.cluster_reads(one.bases, one.offsets)
    assert got.n_clusters == 1 and list(got.mem_id) == [0]


def _digest(cl):
    import hashlib
    h = hashlib.sha256()
    for a in (cl.main_id, cl.main_rev, cl.cl_off, cl.mem_id, cl.mem_rev):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


@pytest.mark.parametrize("genes", [400, 2000])
def test_cluster_config2_full_size_matches_reference_digest(ctx, genes):
    """BASELINE.json configs[1] at full size (2000 genes = 100 k reads): sha256 of the flat cluster set equals the
    digest of the UNMODIFIED reference's output (tests/golden/config2_<genes>.json, made by
    tests/golden/make_golden_config2.py from oracle/_ref; 92 s on 8 CPU threads at 100 k reads)."""
    import hashlib
    import json
    import os
    path = os.path.join(os.path.dirname(__file__), "golden", "config2_%d.json" % genes)
    if not os.path.exists(path):
        pytest.skip("golden not generated")
    gold = json.load(open(path))
    rs = synth.config2(n_genes=genes).sorted_by_length()[0]
    assert hashlib.sha256(rs.bases.tobytes() + rs.offsets.tobytes()).hexdigest() == gold["input_sha256"]
    cl = ctx.cluster_reads(rs.bases, rs.offsets, is_rna=False)
    assert cl.n_clusters == gold["n_clusters"]
    assert _digest(cl) == gold["sha256"]
    # size-independent properties: a partition of the reads, every main_seq is a member of its cluster
    assert sorted(cl.mem_id.tolist()) == list(range(rs.n))
    for c in range(0, cl.n_clusters, 37):
        assert cl.main_id[c] in cl.mem_id[cl.cl_off[c]:cl.cl_off[c + 1]]


def test_pair_similarity_scratch_deferral_and_exhaustion(ctx, orc):
    """survivors with more than 1024 matches sort in a global scratch arena; when it is full they are deferred to the
    follow-up launches (scratch empty again) instead of failing the call, and only what is left after those raises"""
    import rattle_b200
    rng = np.random.default_rng(9)
    core = bytes(rng.choice(list(b"ACGT"), size=2300).astype(np.uint8))
    seqs = [core, core[3:], core[:-5], core[7:-2]]
    rs = synth.from_sequences(seqs)
    ctx.upload(rs.bases, rs.offsets)
    km = [orc.extract_kmers(s, 10, True) for s in seqs]

    def run(n_tasks):
        a = [t % 4 for t in range(n_tasks)]
        b = [(t + 1 + t // 4) % 4 for t in range(n_tasks)]
        a, b = zip(*[(x, y) for x, y in zip(a, b) if x != y])
        res = ctx.pair_similarity(list(a), list(b), [0] * len(a), kmer_size=10, is_rna=False)
        return a, b, res

    ctx.set_option("scratch_mb", 1)  # ~60 KB per survivor: about 16 per launch, 3 launches per call
    try:
        a, b, res = run(40)
        assert res["n_common"].min() > 2000  # every task needs the arena
        for t in range(len(a)):
            first, second = orc.common_kmers(km[a[t]][1], km[a[t]][2], km[b[t]][1], km[b[t]][2])
            bases, dist = orc.similarity(first, second, 10)
            assert res["n_common"][t] == len(first) and res["bases"][t] == bases and res["n_dist"][t] == len(dist)
            assert res["accept"][t] == 1
        with pytest.raises(rattle_b200.RattleError) as e:
            run(400)
        assert e.value.code == -3
    finally:
        ctx.set_option("scratch_mb", 1024)


def test_bulk_copy_ring_scan_matches_numpy_and_clusters(ctx, orc):
    """option bv_kernel=2: scans with at most 16 seeds run k_bv_stream (bitvectors streamed through a shared-memory ring
    by cp.async.bulk + mbarrier): same counts / pass flags as numpy popcounts — on consecutive targets (one 4-KB copy per
    slot), scattered ones and a ragged tail, both strands and RNA — and the same clusters with 7-seed waves"""
    rs = small_set(seed=19, genes=18, per=11)
    k = 10
    popc = np.vectorize(lambda w: bin(int(w)).count("1"))
    bvs = [orc.extract_kmers(rs.seq(i), k, True) for i in range(rs.n)]
    bf = np.stack([b[5] for b in bvs]); br = np.stack([b[6] for b in bvs])
    pc = popc(bf).sum(1)
    ctx.set_option("bv_kernel", 2)
    try:
        for is_rna in (False, True):
            ctx.upload(rs.bases, rs.offsets)
            for seeds, targets in ((np.array([3, 50, 121], np.int32), np.arange(rs.n, dtype=np.int32)),
                                   (np.arange(0, 16, dtype=np.int32), np.arange(rs.n - 3, -1, -2, dtype=np.int32)),
                                   (np.array([7], np.int32), np.arange(5, 42, dtype=np.int32))):
                cf, cr, passed = ctx.bv_scan(seeds, targets, 0.30000000000000004, kmer_size=k, is_rna=is_rna)
                for si, s in enumerate(seeds):
                    ecf = popc(bf[s][None, :] & bf[targets]).sum(1)
                    assert np.array_equal(cf[si], ecf)
                    mmax = np.maximum(pc[s], pc[targets]).astype(np.float64)
                    assert np.array_equal(passed[si] & 1, (ecf / mmax >= 0.30000000000000004).astype(np.uint8))
                    if not is_rna:
                        ecr = popc(bf[s][None, :] & br[targets]).sum(1)
                        assert np.array_equal(cr[si], ecr)
                        assert np.array_equal(passed[si] >> 1, (ecr / mmax >= 0.30000000000000004).astype(np.uint8))
        ctx.set_option("wave", 7)
        got = ctx.cluster_reads(rs.bases, rs.offsets, is_rna=False)
        assert_same_clusters(got, orc.cluster_reads(rs.bases, rs.offsets, is_rna=False, n_threads=8))
    finally:
        ctx.set_option("wave", 512)
        ctx.set_option("bv_kernel", 0)


def _iso_segments(rs, gene_cl):
    """the per-gene read sets of `rattle cluster --iso` (main.cpp:281-298): members by id descending, then stably by
    length descending"""
    lens = rs.lengths()
    segs = []
    for c in range(gene_cl.n_clusters):
        mem = np.asarray(gene_cl.mem_id[gene_cl.cl_off[c]:gene_cl.cl_off[c + 1]], dtype=np.int64)
        mem = np.sort(mem)[::-1]
        segs.append(mem[np.argsort(-lens[mem], kind="stable")])
    return segs


@pytest.mark.parametrize("wave", [512, 16])
def test_cluster_reads_batched_equals_per_gene_calls(ctx, orc, wave):
    """rtl_cluster_reads_batched (one pass over every gene's reads, main.cpp:281-324) == one cluster_reads per gene,
    on the GPU and in the oracle; small wave: candidates of one wave lie in one segment, large: they span many"""
    rs = synth.generate(seed=17, n_genes=30, n_isoforms=2, reads_per_tx=9, len_mean=900.0).sorted_by_length()[0]
    gene_cl = ctx.cluster_reads(rs.bases, rs.offsets, is_rna=False)
    segs = _iso_segments(rs, gene_cl)
    segs.insert(3, np.zeros(0, dtype=np.int64))  # an empty segment is an empty read set
    order = np.concatenate(segs)
    sub = rs.take(order)
    seg_off = np.concatenate([[0], np.cumsum([len(s) for s in segs])]).astype(np.uint32)
    kw = dict(kmer_size=11, t_s=0.3, t_v=25.0)  # --iso-kmer-size / --iso-score-threshold / --iso-max-variance defaults
    ctx.set_option("wave", wave)
    try:
        got, seg_cl_off = ctx.cluster_reads_batched(sub.bases, sub.offsets, seg_off, is_rna=False, **kw)
    finally:
        ctx.set_option("wave", 512)
    assert seg_cl_off[0] == 0 and seg_cl_off[-1] == got.n_clusters
    n_iso = 0
    for s, ids in enumerate(segs):
        c0, c1 = int(seg_cl_off[s]), int(seg_cl_off[s + 1])
        if len(ids) == 0:
            assert c0 == c1
            continue
        one = rs.take(ids)
        exp = orc.cluster_reads(one.bases, one.offsets, k=11, t_s=0.3, t_v=25.0, is_rna=False, n_threads=4)
        assert c1 - c0 == exp["n_clusters"], s
        assert np.array_equal(got.main_id[c0:c1], exp["main_id"]) and np.array_equal(got.main_rev[c0:c1], exp["main_rev"]), s
        m0, m1 = int(got.cl_off[c0]), int(got.cl_off[c1])
        assert np.array_equal(got.cl_off[c0:c1 + 1] - m0, exp["cl_off"]), s
        assert np.array_equal(got.mem_id[m0:m1], exp["mem_id"]) and np.array_equal(got.mem_rev[m0:m1], exp["mem_rev"]), s
        n_iso += c1 - c0
    assert n_iso > gene_cl.n_clusters  # the isoform level splits genes
    # and the per-gene GPU calls give the same
    for s in (0, 5, len(segs) - 1):
        one = rs.take(segs[s])
        single = ctx.cluster_reads(one.bases, one.offsets, is_rna=False, **kw)
        c0, c1 = int(seg_cl_off[s]), int(seg_cl_off[s + 1])
        assert single.n_clusters == c1 - c0 and np.array_equal(single.main_id, got.main_id[c0:c1])


@pytest.mark.parametrize("n", [1, 5, 2047, 2049, 70000])
def test_sort_reads_by_length_is_the_reference_visitation_order(ctx, n):
    """rtl_sort_reads_by_length == sort_read_set (fasta.cpp:458-464): stable, longest first; many equal lengths"""
    rng = np.random.default_rng(n)
    lens = rng.integers(7, 60, size=n) if n < 70000 else rng.integers(150, 400, size=n)
    offsets = np.zeros(n + 1, np.uint64)
    offsets[1:] = np.cumsum(lens)
    got = ctx.sort_by_length(offsets)
    assert np.array_equal(got, np.argsort(-lens.astype(np.int64), kind="stable").astype(np.uint32))
