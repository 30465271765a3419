"""The drop-in CLI's HOST logic in the CPU suite: integration/rattle_dropin.cpp (flattening, the recognition of
main.cpp's per-gene `--iso` loop behind the unmodified cluster_reads signature, the cache of batched results, dealing
genes to device slots) linked with the reference's unmodified main.cpp / fasta.cpp / utils.cpp objects and a MOCK
librattle_b200 that answers rtl_cluster_reads / rtl_cluster_reads_batched from the CPU oracle
(tests/native/mock_rattle_b200.cpp — test infrastructure, nothing of it ships).  clusters.out must equal the reference
CLI's golden digests on real Nanopore reads (tests/golden/cli_toyset.json) whether `--iso` is answered from one batched
call, from per-gene calls, or from two device slots.  The GPU tests (tests/test_cli_gpu.py) run the same flows on the
real library."""
import importlib.util
import json
import os
import subprocess

import pytest

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BUILD = os.path.join(ROOT, "integration", "_build")

_spec = importlib.util.spec_from_file_location("make_golden_cli", os.path.join(HERE, "golden", "make_golden_cli.py"))
gold = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gold)


@pytest.fixture(scope="module")
def mock_cli(tmp_path_factory):
    objs = [os.path.join(BUILD, f) for f in ("main.o", "fasta.o", "utils.o")]
    if not all(os.path.exists(o) for o in objs):
        pytest.skip("integration/_build objects are built where /root/reference exists (__graft_entry__.build())")
    oracle.build(with_ref=False)
    d = tmp_path_factory.mktemp("mockcli")
    ref = os.environ.get("RATTLE_REFERENCE", "/root/reference")
    if not os.path.exists(os.path.join(ref, "cluster.hpp")):
        pytest.skip("reference headers not available")
    dropin_o = str(d / "rattle_dropin.o")
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-I" + os.path.join(ref, "spoa", "include"), "-I" + ref,
                           "-I" + os.path.join(ROOT, "include"), "-c", os.path.join(ROOT, "integration", "rattle_dropin.cpp"),
                           "-o", dropin_o])
    exe = str(d / "rattle_mock")
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-pthread", "-o", exe] + objs +
                          [dropin_o, os.path.join(HERE, "native", "mock_rattle_b200.cpp"), "-L" + os.path.join(ROOT, "oracle"),
                           "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-lz"])
    return exe


def run(exe, fq, out, *flags, env=None):
    os.makedirs(out, exist_ok=True)
    p = subprocess.run([exe, "cluster", "-i", fq, "-o", out, "-t", "4"] + list(flags), check=True, stdout=subprocess.DEVNULL,
                       stderr=subprocess.PIPE, env=dict(os.environ, MOCK_RTL_TRACE="1", RATTLE_B200_TRACE="1", **(env or {})))
    return gold.sha(os.path.join(out, "clusters.out")), p.stderr.decode()


def test_dropin_iso_loop_is_answered_from_one_batched_call(mock_cli, tmp_path):
    want = json.load(open(os.path.join(HERE, "golden", "cli_toyset.json")))["digests"]
    fq = gold.unpack_fixture(str(tmp_path))
    # gene level only: one direct call
    sha, err = run(mock_cli, fq, str(tmp_path / "g"), "--rna")
    assert sha == want["cluster_rna"]["clusters.out"]
    assert err.count("mock: rtl_cluster_reads #") == 1 and "batched" not in err
    # --iso: the gene-level call, then ONE batched call that answers every per-gene call of main.cpp:300
    sha, err = run(mock_cli, fq, str(tmp_path / "i"), "--rna", "--iso")
    assert sha == want["cluster_rna_iso"]["clusters.out"]
    assert err.count("mock: rtl_cluster_reads #") == 1 and err.count("mock: rtl_cluster_reads_batched #") == 1
    assert "clustering directly" not in err
    # the same through per-gene calls (batching switched off): hundreds of direct calls, same file
    sha, err = run(mock_cli, fq, str(tmp_path / "p"), "--rna", "--iso", env={"RATTLE_B200_NO_ISO_BATCH": "1"})
    assert sha == want["cluster_rna_iso"]["clusters.out"]
    assert err.count("mock: rtl_cluster_reads #") > 50 and "batched" not in err
    # two device slots: the genes are dealt to two contexts, one batched call each
    sha, err = run(mock_cli, fq, str(tmp_path / "s"), "--rna", "--iso", env={"RATTLE_B200_DEVICES": "0,0"})
    assert sha == want["cluster_rna_iso"]["clusters.out"]
    assert err.count("mock: rtl_cluster_reads_batched #") == 2
    # both strands
    sha, err = run(mock_cli, fq, str(tmp_path / "c"))
    assert sha == want["cluster_cdna"]["clusters.out"]
