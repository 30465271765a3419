"""The POA kernels' own SOURCE on the CPU.  tests/native/cuda_emu.h is a small SIMT emulation (one OS thread per CUDA
thread, barrier-based warp collectives, the DPX packed-int16 intrinsics, real volatile shared-memory mailboxes);
tests/native/strip_emu_check.cpp compiles rattle_b200/csrc/poa_strip_kernel.cuh and poa_devgraph.cuh against it and
runs k_poa_graph_fold -> k_poa_strip -> k_poa_strip_traceback for every read of small recorded POA runs: the alignments
must equal the reference's (the unmodified reference where oracle/_ref exists, else the CPU restatement), including a
two-strip case where two warps run as a wavefront.  This is a regression test of the kernel logic without a GPU; the
GPU tests (-m gpu) remain the parity tests proper.  The emulation is test infrastructure only."""
import os
import subprocess

import numpy as np

import oracle
from tools import synth

HERE = os.path.dirname(os.path.abspath(__file__))
NATIVE = os.path.join(HERE, "native")


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


def test_poa_kernels_source_emulated_on_cpu(tmp_path):
    exe = str(tmp_path / "strip_emu_check")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-pthread", "-o", exe, "strip_emu_check.cpp"], cwd=NATIVE)
    cases = [pack(1, 5, 90.0), pack(2, 4, 270.0, p_sub=0.05, p_ins=0.04, p_del=0.04)]
    rs = pack(5, 5, 80.0, p_sub=0.08, p_ins=0.05, p_del=0.05)
    seqs = [rs.seq(i) for i in range(rs.n)]
    seqs[2] = seqs[2].replace(b"T", b"U")
    seqs.insert(3, bytes(np.random.default_rng(1).choice(list(b"ACGT"), size=60).astype(np.uint8)))  # unrelated read
    cases.append(synth.from_sequences(seqs))
    files = []
    for i, c in enumerate(cases):
        p = str(tmp_path / ("run%d.txt" % i))
        dump(c, p)
        files.append(p)
    out = subprocess.run([exe] + files, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "mismatches 0" in out.stdout
