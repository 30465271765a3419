"""Kernel-tuning driver for hot path B: `correct_reads` on BASELINE.json configs[3] shape (clusters x 32 forward reads
x 2 kb, clusters.out written directly as SURVEY.md §8d describes), without the clustering path in front.

    python tools/poa_bench.py [--clusters 1000] [--reads 32] [--len 2000] [--iters 2] [--opt key=value ...]

Prints wall time, DP cells, GCUPS of the POA kernels (device time) and of the whole call.  Used under ncu for the
launch list / full-set captures of the POA kernel (profiles/), and for A/B runs (--opt poa_kernel=1 = int32 kernel).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tools import synth  # noqa: E402


def clusters_of(sizes):
    from rattle_b200 import ClusterSet
    off = np.zeros(len(sizes) + 1, np.int64)
    off[1:] = np.cumsum(sizes)
    n = int(off[-1])
    ids = np.arange(n, dtype=np.int32)
    return ClusterSet(ids[off[:-1]].copy(), np.zeros(len(sizes), np.uint8), off, ids, np.zeros(n, np.uint8))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clusters", type=int, default=1000)
    ap.add_argument("--reads", type=int, default=32)
    ap.add_argument("--len", type=int, default=2000)
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--opt", action="append", default=[])
    ap.add_argument("--check", action="store_true", help="compare with the unmodified reference (slow)")
    args = ap.parse_args()
    import rattle_b200
    rs = synth.generate(seed=42, n_genes=args.clusters, n_isoforms=1, reads_per_tx=args.reads, len_mean=float(args.len),
                        len_sd=0.0, len_min=args.len, len_max=args.len, p_flip=0.0, shuffle=False)
    # cluster c = reads [c*R, (c+1)*R) in length-descending order
    lens = rs.lengths()
    order = []
    for c in range(args.clusters):
        idx = np.arange(c * args.reads, (c + 1) * args.reads)
        order.append(idx[np.argsort(-lens[idx], kind="stable")])
    rs = rs.take(np.concatenate(order))
    cl = clusters_of([args.reads] * args.clusters)
    ctx = rattle_b200.Context(0)
    for kv in args.opt:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    out = None
    for it in range(args.iters):
        t0 = time.perf_counter()
        out = ctx.correct_reads(rs.bases, rs.quals, rs.offsets, cl, min_reads=5)
        dt = time.perf_counter() - t0
        st = ctx.stats()
        print(json.dumps({"iter": it, "wall_s": dt, "reads_per_s": rs.n / dt, "poa_cells": st["poa_cells"],
                          "poa_alignments": st["poa_alignments"], "poa_launches": st["poa_launches"],
                          "poa_kernel_ms": st["poa_ms"], "gcups_kernel": st["poa_cells"] / (st["poa_ms"] * 1e6),
                          "gcups_wall": st["poa_cells"] / (dt * 1e9)}))
    if args.check:
        import oracle
        exp = oracle.reference().correct_reads(rs.bases, rs.quals, rs.offsets, cl.as_dict(), min_reads=5,
                                               n_threads=os.cpu_count())
        assert out[2] == exp[2] and out[1] == exp[1], "differs from the reference"
        assert sorted(out[0].split(b"\n")) == sorted(exp[0].split(b"\n"))
        print("parity with the unmodified reference: ok")


if __name__ == "__main__":
    main()
