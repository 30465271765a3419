"""One-off runs of BASELINE.json's other configurations on ONE B200, through the C ABI (profiles/configs_r02.jsonl):

  config 3   1 M cDNA reads (10 k genes x 2 isoforms x 50), `cluster --iso`: gene-level cluster_reads (k=10), then every
             gene's isoform-level clustering (k=11, t_s=0.3, t_v=25: main.cpp:281-324) in ONE rtl_cluster_reads_batched
             pass; for comparison the per-gene loop (one rtl_cluster_reads call per gene) on a sample of genes
  config 4   `correct`: 10 k clusters x 32 reads x 2 kb, clusters.out written directly
  config 5   direct-RNA shape (forward strand, 4/3/3 % errors): cluster --rna + correct on --genes5 genes x 2 x 50 reads

    python tools/config_runs.py [--genes3 10000] [--clusters4 10000] [--genes5 2500]

BASELINE.json shards configs 3-5 over 8 GPUs; the units of work are independent there (genes, clusters), so one GPU's
time for the whole job is what 8 GPUs divide.  Prints one JSON line per configuration.
"""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools import synth  # noqa: E402

ISO_KW = dict(kmer_size=11, t_s=0.3, t_v=25.0)  # --iso-kmer-size / --iso-score-threshold / --iso-max-variance defaults


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()[:16]


def iso_segments(rs, cl):
    """the per-gene read sets of main.cpp:281-298: members by id descending, then stably by length descending"""
    lens = rs.lengths()
    segs = []
    for c in range(cl.n_clusters):
        mem = np.sort(np.asarray(cl.mem_id[cl.cl_off[c]:cl.cl_off[c + 1]], dtype=np.int64))[::-1]
        segs.append(mem[np.argsort(-lens[mem], kind="stable")])
    return segs


def config3(ctx, genes, sample):
    t0 = time.perf_counter()
    rs = synth.config3(n_genes=genes)
    rs = rs.take(ctx.sort_by_length(rs.offsets))  # sort_read_set on the GPU (fasta.cpp:458-464)
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    gene_cl = ctx.cluster_reads(rs.bases, rs.offsets, is_rna=False)
    t_gene = time.perf_counter() - t0
    st_gene = ctx.stats()
    segs = iso_segments(rs, gene_cl)
    sub = rs.take(np.concatenate(segs))
    seg_off = np.concatenate([[0], np.cumsum([len(s) for s in segs])]).astype(np.uint32)
    t0 = time.perf_counter()
    iso, seg_cl_off = ctx.cluster_reads_batched(sub.bases, sub.offsets, seg_off, is_rna=False, **ISO_KW)
    t_iso = time.perf_counter() - t0
    st_iso = ctx.stats()
    # the per-gene loop on a sample of genes (what the drop-in did before batching), checked against the batched result
    pick = np.linspace(0, len(segs) - 1, min(sample, len(segs))).astype(int)
    t0 = time.perf_counter()
    same = True
    for g in pick:
        one = rs.take(segs[g])
        r = ctx.cluster_reads(one.bases, one.offsets, is_rna=False, **ISO_KW)
        c0, c1 = int(seg_cl_off[g]), int(seg_cl_off[g + 1])
        same &= r.n_clusters == c1 - c0 and np.array_equal(r.main_id, iso.main_id[c0:c1])
    t_loop = time.perf_counter() - t0
    per_gene = t_loop / len(pick)
    return {"config": "3: 1 M cDNA reads, cluster --iso (k=10 genes, then k=11 isoforms)", "n_reads": int(rs.n),
            "gene_clusters": int(gene_cl.n_clusters), "isoform_clusters": int(iso.n_clusters),
            "gene_level_s": t_gene, "isoform_level_batched_s": t_iso, "reads_per_s": rs.n / (t_gene + t_iso),
            "per_gene_loop_s_per_gene": per_gene, "per_gene_loop_extrapolated_s": per_gene * len(segs),
            "batched_speedup_over_per_gene_loop": per_gene * len(segs) / t_iso, "sample_genes": int(len(pick)),
            "sample_equals_batched": bool(same), "waves_gene": st_gene["waves"], "waves_iso": st_iso["waves"],
            "kernel_launches_iso": st_iso["kernel_launches"], "digest": digest(iso.main_id, iso.cl_off, iso.mem_id, iso.mem_rev),
            "generate_s": t_gen, "gpus": 1}


def clusters_of(sizes):
    from rattle_b200 import ClusterSet
    off = np.zeros(len(sizes) + 1, np.int64)
    off[1:] = np.cumsum(sizes)
    n = int(off[-1])
    ids = np.arange(n, dtype=np.int32)
    return ClusterSet(ids[off[:-1]].copy(), np.zeros(len(sizes), np.uint8), off, ids, np.zeros(n, np.uint8))


def config4(ctx, clusters):
    rs = synth.config4(n_clusters=clusters)
    lens = rs.lengths()
    order = np.concatenate([np.arange(c * 32, (c + 1) * 32)[np.argsort(-lens[c * 32:(c + 1) * 32], kind="stable")]
                            for c in range(clusters)])
    rs = rs.take(order)
    cl = clusters_of([32] * clusters)
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        out = ctx.correct_reads(rs.bases, rs.quals, rs.offsets, cl, min_reads=5, as_bytes=False)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    st = ctx.stats()
    return {"config": "4: correct, clusters x 32 reads x 2 kb", "clusters": clusters, "n_reads": int(rs.n), "seconds": best,
            "reads_per_s": rs.n / best, "clusters_per_s": clusters / best, "poa_cells": st["poa_cells"],
            "gcups_wall": st["poa_cells"] / best / 1e9, "gcups_busy": st["poa_cells"] / (st["poa_busy_ms"] * 1e6),
            "poa_launches": st["poa_launches"], "consensi_sha256_16": hashlib.sha256(bytes(out[2])).hexdigest()[:16],
            "consensi_records": bytes(out[2]).count(b"\n") // 4, "gpus": 1}


def config5(ctx, genes):
    rs = synth.config5(n_genes=genes).sorted_by_length()[0]
    t0 = time.perf_counter()
    cl = ctx.cluster_reads(rs.bases, rs.offsets, is_rna=True)
    t_cl = time.perf_counter() - t0
    t0 = time.perf_counter()
    out = ctx.correct_reads(rs.bases, rs.quals, rs.offsets, cl, min_reads=5, as_bytes=False)
    t_co = time.perf_counter() - t0
    st = ctx.stats()
    return {"config": "5 (cluster --rna + correct; polish not included): direct-RNA shape, forward strand, 4/3/3 % errors",
            "n_reads": int(rs.n), "clusters": int(cl.n_clusters), "cluster_s": t_cl, "correct_s": t_co,
            "reads_per_s": rs.n / (t_cl + t_co), "gcups_busy": st["poa_cells"] / (st["poa_busy_ms"] * 1e6),
            "consensi_records": bytes(out[2]).count(b"\n") // 4, "gpus": 1}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genes3", type=int, default=10000)
    ap.add_argument("--sample3", type=int, default=150)
    ap.add_argument("--clusters4", type=int, default=10000)
    ap.add_argument("--genes5", type=int, default=2500)
    ap.add_argument("--only", default="3,4,5")
    args = ap.parse_args()
    import rattle_b200
    ctx = rattle_b200.Context(0)
    only = set(args.only.split(","))
    if "4" in only:
        print(json.dumps(config4(ctx, args.clusters4)), flush=True)
    if "3" in only:
        print(json.dumps(config3(ctx, args.genes3, args.sample3)), flush=True)
    if "5" in only:
        print(json.dumps(config5(ctx, args.genes5)), flush=True)


if __name__ == "__main__":
    main()
