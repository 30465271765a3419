"""Larger parity cases against golden digests of the UNMODIFIED reference (tests/golden/make_golden_big.py, made in the
build container from oracle/_ref): cluster + correct on the bench generator (configs[1] shape, 20 k reads, and the
full 100 k-read bench workload when its golden exists), correct on BASELINE.json configs[3] shape (200 clusters x 32
reads x 2 kb), and the reference's WHOLE toy data set through the drop-in CLI (the md5s SURVEY.md 8(c) quotes).
Bar: consensi.fq / uncorrected.fq byte for byte, corrected.fq as a multiset (its order depends on the reference's -t).
"""
import hashlib
import importlib.util
import json
import os
import tempfile

import numpy as np
import pytest

from tools import synth

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
ROOT = os.path.dirname(HERE)


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(GOLD, name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


big = _load("make_golden_big")
cli = _load("make_golden_cli")


def _gold(name):
    path = os.path.join(GOLD, name)
    if not os.path.exists(path):
        pytest.skip("golden %s not generated" % name)
    return json.load(open(path))


@pytest.mark.parametrize("genes", [400, 2000])
def test_config2_cluster_and_correct_match_reference_digests(ctx, genes):
    """bench workload: clusters from the GPU path (equal to the reference's, test_cluster_gpu.py) corrected on the GPU"""
    gold = _gold("config2_%d_correct.json" % genes)
    rs = synth.config2(n_genes=genes).sorted_by_length()[0]
    cl = ctx.cluster_reads(rs.bases, rs.offsets, is_rna=False)
    assert cl.n_clusters == gold["n_clusters"]
    out = ctx.correct_reads(rs.bases, rs.quals, rs.offsets, cl, min_occ=0.3, gap_occ=0.3, err_ratio=30.0, split=200,
                            min_reads=5)
    got = big.fq_digests(out)
    assert got["bytes"] == gold["digests"]["bytes"]
    assert got == gold["digests"]
    st = ctx.stats()
    assert st["poa_cells"] > 0 and st["kernel_launches"] > 0


def test_config4_correct_matches_reference_digests(ctx):
    """BASELINE.json configs[3] shape: clusters x 32 forward reads x 2 kb, clusters.out written directly"""
    from rattle_b200 import ClusterSet
    gold = _gold("config4_200.json")
    rs, cl = big.config4_set(gold["clusters"])
    assert hashlib.sha256(rs.bases.tobytes() + rs.quals.tobytes() + rs.offsets.tobytes()).hexdigest() == gold["input_sha256"]
    cs = ClusterSet(cl["main_id"], cl["main_rev"], cl["cl_off"], cl["mem_id"], cl["mem_rev"])
    out = ctx.correct_reads(rs.bases, rs.quals, rs.offsets, cs, min_occ=0.3, gap_occ=0.3, err_ratio=30.0, split=200,
                            min_reads=5)
    assert big.fq_digests(out) == gold["digests"]
    # size-independent property: one consensus per cluster, in cluster order
    heads = [l for l in out[2].split(b"\n")[0::4] if l]
    assert heads == [b"@gene_cluster_%d reads=32 labels=" % c for c in range(gold["clusters"])]


def test_cli_full_toyset_matches_reference_md5s():
    """BASELINE.json configs[0] and the rest of SURVEY.md 8(c)'s pins on all 8306 reads of the reference's toy set:
    cluster --rna (546 clusters), correct, polish, cluster --rna --iso (939 clusters), cluster (cDNA, both strands)"""
    dropin = os.path.join(ROOT, "integration", "_build", "rattle")
    if not os.path.exists(dropin):
        pytest.fail("integration/_build/rattle is missing: run __graft_entry__.build() where /root/reference exists")
    gold = _gold("cli_toyset_full.json")
    fix = os.path.join(GOLD, gold["fixture"])
    import gzip
    import shutil
    with tempfile.TemporaryDirectory() as wd:
        fastq = os.path.join(wd, "toy.fastq")
        with gzip.open(fix, "rb") as src, open(fastq, "wb") as dst:
            shutil.copyfileobj(src, dst)
        got = cli.run_pipeline(dropin, fastq, wd)
        md5 = hashlib.md5(open(os.path.join(wd, "cluster_rna", "clusters.out"), "rb").read()).hexdigest()
    assert got == gold["digests"]
    assert md5 == "9de962acb7bde7fd555bc2fc3828c2c0"  # SURVEY.md 8(c), BASELINE.json configs[0]
    # the fixture exercises a reverse-strand join: cDNA clustering differs from --rna
    assert got["cluster_cdna"] != got["cluster_rna"]
