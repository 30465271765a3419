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
