"""The drop-in CLI (integration/_build/rattle: the reference's UNMODIFIED main.cpp / fasta.cpp / utils.cpp linked with
integration/rattle_dropin.cpp + librattle_b200 instead of cluster.cpp / correct.cpp / spoa) against

  * committed golden digests of the reference CLI on real Nanopore reads (first 1500 records of the reference's toy
    data set; tests/golden/make_golden_cli.py): cluster --rna, cluster --rna --iso, cluster (cDNA), correct, polish;
  * the reference CLI itself (oracle/_ref/rattle, travels with the repository) on a synthetic both-strand read set,
    including `correct -l` file labels.

Bar: every output file byte for byte; corrected.fq as a multiset of records (its order depends on the reference's -t).
"""
import importlib.util
import json
import os
import subprocess
import tempfile

import pytest

from tools import synth

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
DROPIN = os.path.join(ROOT, "integration", "_build", "rattle")
REF = os.path.join(ROOT, "oracle", "_ref", "rattle")

_spec = importlib.util.spec_from_file_location("make_golden_cli", os.path.join(HERE, "golden", "make_golden_cli.py"))
gold = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gold)


def need_dropin():
    if not os.path.exists(DROPIN):
        pytest.fail("integration/_build/rattle is missing: run __graft_entry__.build() where /root/reference exists")


def test_cli_toyset_matches_reference_digests():
    need_dropin()
    want = json.load(open(os.path.join(HERE, "golden", "cli_toyset.json")))["digests"]
    with tempfile.TemporaryDirectory() as wd:
        got = gold.run_pipeline(DROPIN, gold.unpack_fixture(wd), wd)
    assert got == want


def run(binary, *argv):
    subprocess.run([binary] + list(argv), check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def test_cli_synthetic_both_strands_and_labels_match_reference_binary():
    need_dropin()
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/rattle not built")
    rs = synth.generate(seed=17, n_genes=10, reads_per_tx=24, len_mean=650.0, len_sd=60.0, len_min=300, len_max=1200,
                        p_flip=0.5)
    half = rs.n // 2
    with tempfile.TemporaryDirectory() as wd:
        fa, fb = os.path.join(wd, "a.fastq"), os.path.join(wd, "b.fastq")
        synth.write_fastq(rs.take(range(half)), fa, prefix="a")
        synth.write_fastq(rs.take(range(half, rs.n)), fb, prefix="b")
        outs = {}
        for name, binary in (("ref", REF), ("ours", DROPIN)):
            d = os.path.join(wd, name)
            os.makedirs(d)
            run(binary, "cluster", "-i", fa + "," + fb, "-l", "sa,sb", "-o", d, "-t", "4")
            run(binary, "correct", "-i", fa + "," + fb, "-l", "sa,sb", "-c", os.path.join(d, "clusters.out"), "-o", d,
                "-t", "1", "-r", "3")
            run(binary, "polish", "-i", os.path.join(d, "consensi.fq"), "-o", d, "-t", "4")
            outs[name] = {f: gold.sha(os.path.join(d, f), as_multiset=(f == "corrected.fq"))
                          for f in ("clusters.out", "consensi.fq", "uncorrected.fq", "corrected.fq", "transcriptome.fq")}
            if name == "ours":
                assert b"labels=sa:" in open(os.path.join(d, "consensi.fq"), "rb").read()
        assert outs["ours"] == outs["ref"]


def test_cli_iso_runs_one_batched_pass_and_matches_reference():
    """`cluster --iso` (main.cpp:281-324 calls cluster_reads once per gene): the drop-in answers the per-gene calls from ONE
    rtl_cluster_reads_batched pass; clusters.out equals the reference's, and equals the per-gene path's"""
    need_dropin()
    want = json.load(open(os.path.join(HERE, "golden", "cli_toyset.json")))["digests"]["cluster_rna_iso"]["clusters.out"]
    with tempfile.TemporaryDirectory() as wd:
        fq = gold.unpack_fixture(wd)
        digests = {}
        for name, env in (("batched", {"RATTLE_B200_TRACE": "1"}), ("per_gene", {"RATTLE_B200_NO_ISO_BATCH": "1"})):
            d = os.path.join(wd, name)
            os.makedirs(d)
            p = subprocess.run([DROPIN, "cluster", "-i", fq, "-o", d, "--rna", "--iso", "-t", "4"], check=True,
                               stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, env=dict(os.environ, **env))
            if name == "batched":
                assert b"in one batched pass" in p.stderr and b"clustering directly" not in p.stderr, p.stderr[-2000:]
            digests[name] = gold.sha(os.path.join(d, "clusters.out"))
        assert digests["batched"] == want and digests["per_gene"] == want


def test_cli_several_device_slots_give_the_same_files():
    """RATTLE_B200_DEVICES with two slots (both on GPU 0 here): genes of --iso and clusters of `correct` are split over the
    slots and merged; clusters.out / consensi.fq byte-identical, corrected / uncorrected as multisets of records"""
    need_dropin()
    want = json.load(open(os.path.join(HERE, "golden", "cli_toyset.json")))["digests"]
    env = dict(os.environ, RATTLE_B200_DEVICES="0,0", RATTLE_B200_TRACE="1")
    with tempfile.TemporaryDirectory() as wd:
        fq = gold.unpack_fixture(wd)
        d = os.path.join(wd, "o")
        os.makedirs(d)
        p = subprocess.run([DROPIN, "cluster", "-i", fq, "-o", d, "--rna", "--iso", "-t", "4"], check=True,
                           stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, env=env)
        assert p.stderr.count(b"in one batched pass") == 2, p.stderr[-2000:]
        assert gold.sha(os.path.join(d, "clusters.out")) == want["cluster_rna_iso"]["clusters.out"]
        subprocess.run([DROPIN, "cluster", "-i", fq, "-o", d, "--rna", "-t", "4"], check=True, stdout=subprocess.DEVNULL,
                       stderr=subprocess.DEVNULL, env=env)
        p = subprocess.run([DROPIN, "correct", "-i", fq, "-c", os.path.join(d, "clusters.out"), "-o", d, "-t", "1"], check=True,
                           stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, env=env)
        assert p.stderr.count(b"corrected") >= 2, p.stderr[-2000:]
        assert gold.sha(os.path.join(d, "consensi.fq")) == want["correct"]["consensi.fq"]
        assert gold.sha(os.path.join(d, "corrected.fq"), as_multiset=True) == want["correct"]["corrected.fq"]
        # uncorrected.fq: same records; their order depends on how the clusters were dealt to the slots
        with tempfile.TemporaryDirectory() as wd1:
            d1 = os.path.join(wd1, "one")
            os.makedirs(d1)
            subprocess.run([DROPIN, "correct", "-i", fq, "-c", os.path.join(d, "clusters.out"), "-o", d1, "-t", "1"], check=True,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            assert gold.sha(os.path.join(d1, "uncorrected.fq")) == want["correct"]["uncorrected.fq"]
            assert gold.sha(os.path.join(d, "uncorrected.fq"), as_multiset=True) == gold.sha(os.path.join(d1, "uncorrected.fq"), as_multiset=True)
