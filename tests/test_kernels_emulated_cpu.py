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


def dump_with_rows(rs, path):
    lib = oracle.reference() if oracle.have_ref() else oracle.oracle()
    rows, alns = lib.poa_msa(rs.bases, rs.offsets, want_alignments=True)
    dump(rs, path)
    with open(path, "a") as f:
        for r in rows:
            f.write(r.decode() + "\n")


def test_device_resident_poa_chain_emulated_on_cpu(tmp_path):
    """poa_devchain.cuh: the whole per-pack loop (warp-cooperative Graph::add_alignment, sort, row records, DP,
    traceback, MSA) with no host graph in between; alignments and MSA rows must equal the reference's.  A second run with
    shrunken capacities checks that a pack that outgrows its slot is flagged (and nothing else breaks)."""
    exe = str(tmp_path / "chain_emu_check")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-pthread", "-o", exe, "chain_emu_check.cpp"], cwd=NATIVE)
    cases = [pack(1, 6, 90.0), pack(2, 5, 270.0, p_sub=0.05, p_ins=0.04, p_del=0.04),
             pack(11, 9, 120.0, p_sub=0.10, p_ins=0.06, p_del=0.06)]
    rs = pack(5, 6, 80.0, p_sub=0.08, p_ins=0.05, p_del=0.05)
    seqs = [rs.seq(i) for i in range(rs.n)]
    seqs[2] = seqs[2].replace(b"T", b"U")
    seqs.insert(3, bytes(np.random.default_rng(1).choice(list(b"ACGT"), size=60).astype(np.uint8)))  # unrelated read
    seqs.append(seqs[0][20:50])  # aligned in the middle of the graph: nothing before / after the aligned part is new
    seqs.append(b"GGGG" + seqs[1] + b"CCCCC")  # new prefix and suffix chains
    cases.append(synth.from_sequences(seqs))
    files = []
    for i, c in enumerate(cases):
        p = str(tmp_path / ("run%d.txt" % i))
        dump_with_rows(c, p)
        files.append(p)
    for smem in ("1", "0"):  # topological sort out of shared memory / in global memory (oversized graphs)
        out = subprocess.run([exe, "100", smem] + files, capture_output=True, text=True, timeout=900)
        assert out.returncode == 0, out.stdout + out.stderr
        assert "failed packs 0, msa mismatches 0" in out.stdout
    out = subprocess.run([exe, "30", "1"] + files, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "failed packs 0" not in out.stdout and "msa mismatches 0" in out.stdout


def test_parallel_topological_sort_emulated_on_cpu(tmp_path):
    """dc_sort_blocks (label propagation + per-block DFS by a whole CTA) == the reference's sort after every
    add_alignment, on graphs with long branches (serial redo of oversized blocks), bushy bubbles and big aligned groups"""
    exe = str(tmp_path / "sort_emu_check")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-pthread", "-o", exe, "sort_emu_check.cpp"], cwd=NATIVE)
    cases = [pack(3, 16, 500.0), pack(21, 30, 300.0, p_sub=0.08, p_ins=0.06, p_del=0.06)]
    rs = pack(31, 10, 600.0)
    seqs = [rs.seq(i) for i in range(rs.n)]
    rng = np.random.default_rng(0)
    unrelated = bytes(rng.choice(list(b"ACGT"), size=250).astype(np.uint8))
    seqs.insert(2, unrelated)                                   # a separate chain
    seqs.insert(5, seqs[0][:200] + unrelated[:120] + seqs[0][200:])  # a 120-node branch that joins back
    seqs.insert(7, unrelated + seqs[1][300:])                   # joins the unrelated chain to the backbone
    seqs.insert(9, b"G" * 40)
    seqs.append(seqs[3].replace(b"T", b"U"))                    # aligned groups with a fifth letter
    cases.append(synth.from_sequences(seqs))
    files = []
    for i, c in enumerate(cases):
        p = str(tmp_path / ("run%d.txt" % i))
        dump(c, p)
        files.append(p)
    for threads in ("256", "64"):
        out = subprocess.run([exe, threads] + files, capture_output=True, text=True, timeout=900)
        assert out.returncode == 0, out.stdout + out.stderr
        assert "order mismatches 0, lead mismatches 0" in out.stdout
