"""The product's MSA end clean-up (rattle_b200/csrc/msa_ends.hpp, a restatement of correct.cpp:32-92 as two
trim-the-front passes) against the reference's own fix_msa_ends (oracle/_ref/libref_shim.so: ref_fix_msa_ends forwards
to /root/reference/correct.cpp:32) on random MSA rows built to hit its cases: short leading / trailing blocks followed
by long gap runs, blocks of exactly 9 / 10 bases, gap runs of exactly 19 / 20, rows that are blanked completely (the
reference leaves those reversed when it happens on the second pass), all-gap rows."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))


def random_row(rng, ncol):
    """a row as segments: [gaps][block][gaps][block]...; block and gap lengths drawn around the decision limits"""
    row = []
    while len(row) < ncol:
        row += ["-"] * int(rng.choice([0, 1, 3, 4, 5, 19, 20, 21, 30]))
        blk = int(rng.choice([1, 3, 9, 10, 11, 25]))
        for _ in range(blk):
            row.append(rng.choice(list("ACGT")))
            if rng.random() < 0.15:
                row += ["-"] * int(rng.integers(1, 4))  # short gaps inside a block (fewer than 4: same block)
    row = row[:ncol]
    return "".join(row)


def cases(seed, n_cases):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_cases):
        ncol = int(rng.integers(1, 120))
        n = int(rng.integers(1, 6))
        rows = [random_row(rng, ncol) for _ in range(n)]
        if rng.random() < 0.2:
            rows[0] = "-" * ncol
        if rng.random() < 0.2:
            rows[-1] = "-" * (ncol // 2) + "ACG"[:max(0, min(3, ncol - ncol // 2))] + "-" * max(0, ncol - ncol // 2 - 3)
        rows = [r[:ncol].ljust(ncol, "-") for r in rows]
        seqs = [r.replace("-", "") for r in rows]
        quals = ["".join(chr(int(x)) for x in rng.integers(48, 75, len(s))) for s in seqs]
        out.append((ncol, rows, seqs, quals))
    return out


def reference_fix(lib, ncol, rows, seqs, quals):
    n = len(rows)
    rb = ctypes.create_string_buffer("".join(rows).encode(), n * ncol + 1)
    off = np.zeros(n + 1, np.int64)
    off[1:] = np.cumsum([len(s) for s in seqs])
    sb = ctypes.create_string_buffer("".join(seqs).encode(), int(off[-1]) + 1)
    qb = ctypes.create_string_buffer("".join(quals).encode(), int(off[-1]) + 1)
    new_len = np.zeros(n, np.int32)
    lib.ref_fix_msa_ends.restype = ctypes.c_int
    lib.ref_fix_msa_ends.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p,
                                     ctypes.c_void_p, ctypes.c_void_p]
    lib.ref_fix_msa_ends(rb, n, ncol, sb, qb, off.ctypes.data, new_len.ctypes.data)
    res = []
    for i in range(n):
        res.append((rb.raw[i * ncol:(i + 1) * ncol].decode(), sb.raw[off[i]:off[i] + new_len[i]].decode(),
                    qb.raw[off[i]:off[i] + new_len[i]].decode()))
    return res


def test_msa_end_cleanup_equals_reference(tmp_path):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/libref_shim.so not built")
    lib = ctypes.CDLL(oracle.REF_SHIM_SO)
    if not hasattr(lib, "ref_fix_msa_ends"):
        pytest.skip("libref_shim.so predates ref_fix_msa_ends: rebuild where /root/reference exists")
    exe = str(tmp_path / "msa_ends_check")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", exe, "msa_ends_check.cpp"], cwd=os.path.join(HERE, "native"))
    cs = cases(7, 400)
    text = []
    for ncol, rows, seqs, quals in cs:
        text.append("%d %d" % (len(rows), ncol))
        for r, s, q in zip(rows, seqs, quals):
            text.append("%s %s %s" % (r, s or ".", q or "."))
    out = subprocess.run([exe], input="\n".join(text) + "\n", capture_output=True, text=True, check=True).stdout.split("\n")
    at = 0
    trimmed = reversed_left = 0
    for ncol, rows, seqs, quals in cs:
        exp = reference_fix(lib, ncol, rows, seqs, quals)
        for i, (er, es, eq) in enumerate(exp):
            gr, gs, gq = out[at].split(" ")
            at += 1
            assert (gr, gs if gs != "." else "", gq if gq != "." else "") == (er, es, eq), (rows[i], er, gr)
            trimmed += es != seqs[i] and len(es) < len(seqs[i])
            reversed_left += er.replace("-", "") != "" and er.replace("-", "") == seqs[i][::-1] != seqs[i]
    assert trimmed > 50  # the cases really exercise the trimming
