"""The device-side MSA post-processing (rattle_b200/csrc/poa_vote.cuh — fix_msa_ends, column vote, read correction,
consensus; the bodies of k_vote_rows / k_vote_cols / k_vote_apply / k_vote_consensus compile for the host) against the
host restatement of correct.cpp:32-309 (msa_ends.hpp / vote_host.hpp, which the GPU tests pin against the unmodified
reference), on real POA multiple sequence alignments of noisy packs and on synthetic rows built around the decision limits
(short isolated blocks at the ends, equal qualities -> mean errors that are table entries, ties in the vote, low coverage).
tests/native/vote_check.cpp does the comparison; here the cases are made."""
import os
import subprocess

import numpy as np

import oracle
from tools import synth

HERE = os.path.dirname(os.path.abspath(__file__))
NATIVE = os.path.join(HERE, "native")


def msa_case(rs, rows, min_occ=0.3, gap_occ=0.3):
    ncol = len(rows[0])
    out = ["%d %d %r %r" % (rs.n, ncol, min_occ, gap_occ)]
    for i in range(rs.n):
        out.append("%s %s %s" % (rows[i].decode(), rs.seq(i).decode(), rs.qual(i).decode()))
    return "\n".join(out)


def synthetic_case(rng, n, ncol, qual_mode):
    """random rows with a shared backbone: [short block][long gap] heads/tails, internal gaps, mismatches"""
    back = rng.choice(list("ACGT"), size=ncol)
    lines = []
    for _ in range(n):
        a = int(rng.integers(0, ncol // 3))
        b = int(rng.integers(2 * ncol // 3, ncol))
        row = np.array(["-"] * ncol)
        for k in range(a, b):
            r = rng.random()
            if r < 0.80:
                row[k] = back[k]
            elif r < 0.90:
                row[k] = rng.choice(list("ACGT"))
        if rng.random() < 0.5 and a > 30:  # a short isolated block in front of the read
            s0 = int(rng.integers(0, a - 28))
            blk = int(rng.choice([3, 9, 10]))
            row[s0:s0 + blk] = rng.choice(list("ACGT"), size=blk)
            row[s0 + blk:a] = "-"
            if rng.random() < 0.5:
                row[a:a + 22] = "-"
        if rng.random() < 0.5 and ncol - b > 30:
            blk = int(rng.choice([2, 9, 11]))
            row[ncol - blk:] = rng.choice(list("ACGT"), size=blk)
        row = "".join(row)
        seq = row.replace("-", "")
        if len(seq) == 0:
            row = "A" + row[1:]
            seq = "A"
        if qual_mode == 0:
            q = rng.integers(36, 74, size=len(seq))
        elif qual_mode == 1:
            q = np.full(len(seq), 53)  # every mean error is a table entry (or a sum of equal ones)
        else:
            q = rng.choice([40, 41, 73], size=len(seq))
        lines.append("%s %s %s" % (row, seq, "".join(chr(c) for c in q)))
    return "%d %d %r %r\n%s" % (n, ncol, float(rng.choice([0.3, 0.0, 0.6])), float(rng.choice([0.3, 0.5])), "\n".join(lines))


def test_device_vote_routines_equal_host_restatement(tmp_path):
    exe = str(tmp_path / "vote_check")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", exe, "vote_check.cpp"], cwd=NATIVE)
    lib = oracle.reference() if oracle.have_ref() else oracle.oracle()
    cases = []
    for seed, n, length, kw in ((1, 12, 300.0, {}), (2, 30, 500.0, dict(p_sub=0.06, p_ins=0.04, p_del=0.04)),
                                (3, 7, 200.0, dict(p_sub=0.10, p_ins=0.06, p_del=0.06))):
        rs = synth.generate(seed=seed, n_genes=1, reads_per_tx=n, len_mean=length, len_sd=30.0, len_min=100, len_max=900,
                            p_flip=0.0, shuffle=False, **kw)
        rs = rs.sorted_by_length()[0]
        rows = lib.poa_msa(rs.bases, rs.offsets)
        cases.append(msa_case(rs, rows))
        cases.append(msa_case(rs, rows, min_occ=0.0, gap_occ=0.6))
    rng = np.random.default_rng(7)
    for i in range(60):
        cases.append(synthetic_case(rng, int(rng.integers(1, 9)), int(rng.integers(60, 260)), i % 3))
    out = subprocess.run([exe], input="\n".join(cases) + "\n", capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr
    lines = out.stdout.split()
    assert len(out.stdout.strip().splitlines()) == len(cases)
    assert "DIFF" not in out.stdout
    assert out.stdout.count("ok") >= len(cases) - 10  # (a few synthetic packs may be degenerate: those go to the host)
    assert "mismatches 0" in out.stderr
    # the exactness machinery was exercised: table symbols and host-flagged columns both occurred
    stats = out.stderr.strip().split(",")
    flagged = int(stats[4].split()[-1])
    table = int(stats[5].split()[-1])
    assert table > 0 and flagged >= 0
