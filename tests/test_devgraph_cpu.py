"""poa_devgraph.cuh on the CPU: the routines the GPU runs to mirror, sort and describe the partial-order graphs
(k_poa_graph_fold) are plain functions of raw arrays, so they are compiled for the host here and replayed against
PoaGraph (the host graph: literal restatement of spoa's Graph, graph.cpp:154-353) on recorded POA runs:
rank order identical after every add_alignment, row records equivalent (same predecessor rows, same spilled rows).
The alignments come from the oracle (the unmodified reference where oracle/_ref exists, else the CPU restatement)."""
import os
import subprocess

import numpy as np
import pytest

import oracle
from tools import synth

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "devgraph_check.cpp")


def dump(rs, path):
    lib = oracle.reference() if oracle.have_ref() else oracle.oracle()
    rows, alns = lib.poa_msa(rs.bases, rs.offsets, want_alignments=True)
    with open(path, "w") as f:
        f.write("%d\n" % rs.n)
        for i in range(rs.n):
            f.write(rs.seq(i).decode() + "\n")
        for a in alns:
            f.write("%d\n" % len(a))
            for x, y in a:
                f.write("%d %d\n" % (x, y))


def pack(seed, n, length, **kw):
    return synth.generate(seed=seed, n_genes=1, reads_per_tx=n, len_mean=length, len_sd=0.0, len_min=int(length),
                          len_max=int(length), p_flip=0.0, shuffle=False, **kw)


def test_device_graph_routines_match_host_graph(tmp_path):
    exe = str(tmp_path / "devgraph_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, SRC], cwd=os.path.dirname(SRC))
    files = []
    cases = [pack(3, 20, 500.0), pack(21, 40, 400.0, p_sub=0.06, p_ins=0.04, p_del=0.04)]
    rs = pack(31, 8, 600.0)
    seqs = [rs.seq(i) for i in range(rs.n)]
    seqs[3] = seqs[3][:200]
    seqs.insert(2, bytes(np.random.default_rng(0).choice(list(b"ACGT"), size=250).astype(np.uint8)))  # unrelated read
    seqs.insert(6, b"G" * 40)
    cases.append(synth.from_sequences(seqs))
    for i, c in enumerate(cases):
        p = str(tmp_path / ("run%d.txt" % i))
        dump(c, p)
        files.append(p)
    out = subprocess.run([exe] + files, capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "order mismatches 0, record mismatches 0, deferred-sort msa mismatches 0" in out.stdout
