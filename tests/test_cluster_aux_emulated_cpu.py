"""The clustering path's small kernels on the CPU: tests/native/cluster_aux_emu_check.cpp compiles the SOURCE of
rattle_b200/csrc/cluster_aux_kernels.cuh — k_pack_bases (2-bit read ingest), the bitonic length sort behind
rtl_sort_reads_by_length, and the greedy waves' bookkeeping (k_select, k_mark_cand, k_resolve and its CTA-wide form
k_resolve_cta, k_select_seg of the batched `--iso` clustering, k_apply) — against tests/native/cuda_emu.h (one OS thread per
CUDA thread) and checks every kernel against straightforward host code: packing codes and the bad-base flag, the stable
longest-first order of fasta.cpp:458-464, the greedy resolution of cluster.cpp:124-166 on random decision matrices (both
resolve kernels), first-untaken-item-per-segment selection.  The header is product code with identical SASS on the device;
the emulation is test infrastructure only.  GPU parity of the whole path: tests/test_cluster_gpu.py."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
NATIVE = os.path.join(HERE, "native")


def test_cluster_aux_kernels_source_emulated_on_cpu(tmp_path):
    exe = str(tmp_path / "cluster_aux_emu_check")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-pthread", "-o", exe, "cluster_aux_emu_check.cpp"], cwd=NATIVE)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "failures 0" in out.stdout and "FAIL" not in out.stdout
    for what in ("ok pack", "ok sort 5000", "ok resolve W=1024", "ok select_seg", "ok apply"):
        assert what in out.stdout
