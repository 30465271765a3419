#!/usr/bin/env python
"""bench.py — reads/s of RATTLE's cluster+correct hot path on synthetic cDNA reads (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--genes G]

A "step" is one full pass of the hot path over the workload through the reference-shaped C-ABI calls with HOST
buffers: rtl_cluster_reads (H2D of the reads, k-mer extraction, greedy bitvector/k-mer clustering: initial pass +
merge rounds, D2H of the cluster set) and rtl_correct_reads (POA correction + consensus of the resulting clusters,
FASTQ text written to host buffers).  ONE timed loop gives both numbers:

  e2e    reads / step time (host -> host, copies inside the timed region)
  value  reads / (step time - device time of the H2D copy of the read set), i.e. with the reads already resident in
         HBM; the copy is timed with its own CUDA events inside the library (rtl_stats.upload_ms)

Workload: BASELINE.json configs[1] shape (tools/synth.config2: genes x 50 reads x ~1.5 kb, both strands) at 2500
genes = 125 k reads PER GPU, weak scaling, so that 8 GPUs run the 1 M reads BASELINE.json's metric names
(--genes 2000 = configs[1] itself, 100 k reads).  N>1 (torchrun): every greedy wave's (seed, target) pairs are
sharded over the ranks with one NCCL min-allreduce of the decision arrays per wave phase; clusters are sharded over
the ranks for correction (no collective).  After the timed loop the outputs are digested (cluster set, consensi.fq,
corrected/uncorrected as order-independent multiset digests) and, for N>1, rank 0 re-runs the whole workload
UNSHARDED once (untimed) and the digests must be equal — a sharded run that differs fails loudly.

--impl reference times the UNMODIFIED reference (oracle/_ref/libref_shim.so, compiled from /root/reference) on the
host cores, on a bounded sample of the same generator sized to the box's core count.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")  # before torch creates the CUDA context (rattle_b200/__init__.py)

from tools import synth  # noqa: E402

CLUSTER_KW = dict(kmer_size=10, t_s=0.2, t_v=1e6, bv_threshold=0.4, min_bv_threshold=0.2, bv_falloff=0.05,
                  repr_percentile=0.15, is_rna=False)  # main.cpp:200-221 defaults, cDNA (both strands)
CORRECT_KW = dict(min_occ=0.3, gap_occ=0.3, err_ratio=30.0, split=200, min_reads=5)  # main.cpp:396-402


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows = []
        self.stop = threading.Event()
        self.index = index
        self.t = None

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([x.strip() for x in line.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=10)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def make_workload(genes):
    rs = synth.config2(n_genes=genes)
    return rs.sorted_by_length()[0]  # main.cpp:254 sort_read_set


# ------------------------------------------------------------------------------------------------ output digests
def cluster_digest(cl):
    """sha256 of the flat cluster set (same as tests/golden/make_golden_config2.py: the clusters.out content)"""
    h = hashlib.sha256()
    for a in (cl.main_id, cl.main_rev, cl.cl_off, cl.mem_id, cl.mem_rev):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def records(text):
    lines = bytes(text).split(b"\n")
    return [b"\n".join(lines[i:i + 4]) for i in range(0, len(lines) - 1, 4)]


def multiset_sum(recs):
    """order-independent digest of a multiset of FASTQ records that adds up over shards: sum of sha256(record) mod 2^256"""
    s = 0
    for r in recs:
        s += int.from_bytes(hashlib.sha256(r).digest(), "big")
    return s % (1 << 256)


def correction_digests(parts):
    """parts: per-rank (corrected, uncorrected, consensi) texts -> digests of the whole job's output"""
    from rattle_b200.dist import merge_consensi
    cons = merge_consensi([p[2] for p in parts])
    return {"consensi_sha256": hashlib.sha256(cons).hexdigest(), "consensi_records": cons.count(b"\n") // 4,
            "corrected_sum256": "%064x" % (sum(multiset_sum(records(p[0])) for p in parts) % (1 << 256)),
            "uncorrected_sum256": "%064x" % (sum(multiset_sum(records(p[1])) for p in parts) % (1 << 256))}


def golden_form_digests(out):
    """the digest forms tests/golden/make_golden_big.py stores (N=1 only: needs every record on one rank)"""
    return {"corrected_sorted": hashlib.sha256(b"\n".join(sorted(records(out[0])))).hexdigest(),
            "uncorrected": hashlib.sha256(bytes(out[1])).hexdigest(),
            "consensi": hashlib.sha256(bytes(out[2])).hexdigest()}


# ------------------------------------------------------------------------------------------------ reference arm
def ref_sample_genes(args, steps=None):
    """bounded CPU sample of the reference arm: the whole (warmup + steps) run has to end within a few minutes, so one
    step gets min(20 s, 300 s / (warmup + steps)) of reference work; measured on the bench box: 0.75 genes (x 50
    reads) per core and second, 0.6 taken for margin (correct dominates at these sizes and is linear in reads, clustering ~quadratic)"""
    if args.ref_genes > 0:
        return args.ref_genes
    cores = os.cpu_count() or 1
    n = max(1, steps if steps is not None else args.warmup + args.steps)
    per_step_s = min(20.0, 300.0 / n)
    return int(min(2000, max(40, 0.6 * cores * per_step_s)))


def reference_arm(args, rank, world):
    """Unmodified reference on the host cores (rank 0 only)."""
    if rank != 0:
        return
    import oracle
    cores = os.cpu_count() or 1
    if not oracle.have_ref():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libref_shim.so was not built in the container"}))
        return
    ref = oracle.reference()
    genes = ref_sample_genes(args)
    rs = make_workload(genes)
    sample = ("same generator at %d genes x 50 = %d reads (~1.5 kb cDNA), cluster%s, %d threads; sample sized to the "
              "core count and to warmup+steps so that the whole run is <= ~5 min" % (genes, rs.n, "+correct" if args.correct else "",
                                                                             cores))
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        cl = ref.cluster_reads(rs.bases, rs.offsets, k=10, t_s=0.2, t_v=1e6, bv_thr=0.4, bv_min=0.2, bv_falloff=0.05,
                               repr_pct=0.15, is_rna=False, n_threads=cores)
        if args.correct:
            ref.correct_reads(rs.bases, rs.quals, rs.offsets, cl, n_threads=cores, **CORRECT_KW)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    v = rs.n / (ms / 1e3)
    line = {"impl": "reference", "metric": "reads/sec cluster+correct" if args.correct else "reads/sec cluster",
            "value": v, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPE,
            "data": "synthetic", "config": workload_config(args, rs.n, genes, 1),
            "cpu_baseline": {"value": v, "unit": "reads/s", "cores": cores, "kind": "reference", "sample": sample,
                             "full_config_one_off": full_config_reference()},
            "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def full_config_reference():
    """one-off run of the unmodified reference on the whole configs[1] workload (100 k reads), made in the build
    container by tests/golden/make_golden_big.py (the reference cannot repeat it 25 times inside a bench run)"""
    p = os.path.join(ROOT, "tests", "golden", "config2_2000_correct.json")
    if not os.path.exists(p):
        return None
    g = json.load(open(p))
    s = g["cluster_seconds"] + g["correct_seconds"]
    return {"n_reads": g["n_reads"], "seconds": s, "reads_per_s": g["n_reads"] / s, "threads": g["threads"],
            "where": "build container (tests/golden/config2_2000_correct.json)"}


DTYPE = "int16x2 (POA DP) / u64 popcount + u32 k-mer hash (clustering)"


def workload_config(args, n_reads, genes, world):
    return {"workload": "BASELINE.json configs[1] shape, %d genes x 50 reads per GPU: %d synthetic cDNA reads x ~1.5 kb in "
                        "total (3/2/2%% sub/ins/del, random strand), k=10 gene clustering%s%s" % (
                            genes // max(1, world), n_reads, " + correct" if args.correct else "",
                            "; = the 1 M reads of BASELINE.json's metric" if n_reads == 1000000 else ""),
            "n_reads": int(n_reads), "kmer_size": 10, "strands": 2, "l2": "inputs larger than L2 (k-mer lists + "
            "bitvectors of the workload exceed 126 MB); no explicit flush",
            "parallelism": "reads block-sharded x%d for k-mer extraction + broadcast of the blocks, wave pairs sharded x%d + "
                           "min-allreduce of the decisions (cluster), clusters round-robin x%d (correct)" % (world, world, world)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genes", type=int, default=2500,
                    help="genes PER GPU (x 50 reads): 2500 = 125 k reads per GPU, so that 8 GPUs run the 1 M reads of "
                         "BASELINE.json's metric; 2000 = configs[1] itself (100 k reads)")
    ap.add_argument("--ref-genes", type=int, default=0, help="size of the bounded CPU sample (x50 reads); 0 = 15 per core")
    ap.add_argument("--no-correct", dest="correct", action="store_false", default=True)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-check", action="store_true", help="N>1: skip the unsharded re-run that the digests are compared with")
    ap.add_argument("--opt", action="append", default=[], help="library tunable key=value (rtl_set_option)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import rattle_b200
    from rattle_b200.dist import make_allreduce_callback, make_broadcast_callback, shard_clusters

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    # an explicit stream: the legacy default stream (handle 0) would make the library fall back to its own stream,
    # and torch's NCCL calls would no longer be ordered against the library's kernels
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = rattle_b200.Context(local_rank)
    for kv in args.opt:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    ctx.set_stream(stream.cuda_stream)
    allreduce_cb = make_allreduce_callback(stream.cuda_stream) if world > 1 else None
    if world > 1:
        ctx.set_shard(rank, world, allreduce_cb)
        ctx.set_broadcast(make_broadcast_callback(stream.cuda_stream))

    total_genes = args.genes * world
    rs = make_workload(total_genes)
    n_reads = rs.n
    pin_bases = torch.from_numpy(rs.bases).pin_memory()
    pin_quals = torch.from_numpy(rs.quals).pin_memory()
    bases_np = pin_bases.numpy()
    quals_np = pin_quals.numpy()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def correct_shard(cl, sharded=True):
        """clusters sharded round-robin over ranks (independent packs, no collective); headers carry global ids"""
        if world == 1 or not sharded:
            sub, gids = cl, None
        else:
            sub, gids = shard_clusters(cl, rank, world)
        # results stay in the Context's host buffers (what the C ABI wrote): no copy into Python bytes objects
        return ctx.correct_reads(bases_np, quals_np, rs.offsets, sub, as_bytes=False, cluster_ids=gids, **CORRECT_KW)

    def step():
        """one pass of the hot path, host buffers in, host buffers out"""
        cl = ctx.cluster_reads(bases_np, rs.offsets, **CLUSTER_KW)
        st = ctx.stats()
        out, st2 = None, None
        if args.correct:
            out = correct_shard(cl)
            st2 = ctx.stats()
        return cl, st, st2, out

    for _ in range(args.warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.__enter__()
    res = []
    e0.record(stream)
    for _ in range(args.steps):
        res.append(step())
    e1.record(stream)
    barrier()
    if sampler:
        sampler.__exit__()
    ms_total = e0.elapsed_time(e1)
    upload_ms = sum(r[1]["upload_ms"] for r in res)
    t = torch.tensor([ms_total / args.steps, (ms_total - upload_ms) / args.steps], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e, ms_res = float(t[0].item()), float(t[1].item())
    clocks = sampler.summary() if sampler else None

    cl, st, st2, out = res[-1]
    launches = sum(r[1]["kernel_launches"] + (r[2]["kernel_launches"] if r[2] else 0) for r in res)
    h2d = st["h2d_bytes"] + (st2["h2d_bytes"] if st2 else 0)
    d2h = st["d2h_bytes"] + int(cl.n_clusters) * 13 + n_reads * 5 + (st2["d2h_bytes"] if st2 else 0)

    # ---- digests of the job's output (untimed): every rank holds the same cluster set and its share of the FASTQ texts
    digests = {"clusters_sha256": cluster_digest(cl)}
    mine = tuple(bytes(x) for x in out) if out is not None else None
    if args.correct:
        if world > 1:
            # the corrected reads stay where they are: only their order-independent digest and the consensi travel
            part = (multiset_sum(records(mine[0])), multiset_sum(records(mine[1])), mine[2])
            parts = [None] * world if rank == 0 else None
            dist.gather_object(part, parts, dst=0)
            if rank == 0:
                from rattle_b200.dist import merge_consensi
                cons = merge_consensi([p[2] for p in parts])
                digests.update({"consensi_sha256": hashlib.sha256(cons).hexdigest(), "consensi_records": cons.count(b"\n") // 4,
                                "corrected_sum256": "%064x" % (sum(p[0] for p in parts) % (1 << 256)),
                                "uncorrected_sum256": "%064x" % (sum(p[1] for p in parts) % (1 << 256))})
        else:
            digests.update(correction_digests([mine]))
            digests["golden_form"] = golden_form_digests(mine)
            gp = os.path.join(ROOT, "tests", "golden", "config2_%d_correct.json" % total_genes)
            if os.path.exists(gp):  # the unmodified reference's digests of this very workload
                g = json.load(open(gp))["digests"]
                digests["equals_reference_golden"] = all(digests["golden_form"][k] == g[k] for k in digests["golden_form"])
        gc = os.path.join(ROOT, "tests", "golden", "config2_%d.json" % total_genes)
        if os.path.exists(gc):
            digests["clusters_equal_reference_golden"] = json.load(open(gc))["sha256"] == digests["clusters_sha256"]
    # ---- N>1: the sharded outputs must equal an unsharded run of the same workload (one untimed check step, rank 0)
    if world > 1 and not args.no_check:
        if rank == 0:
            ctx.set_shard(0, 1, None)
            ctx.set_broadcast(None)
            t0 = time.perf_counter()
            cl1 = ctx.cluster_reads(bases_np, rs.offsets, **CLUSTER_KW)
            ref_d = {"clusters_sha256": cluster_digest(cl1)}
            if args.correct:
                o1 = correct_shard(cl1, sharded=False)
                ref_d.update(correction_digests([tuple(bytes(x) for x in o1)]))
            digests["unsharded_check_s"] = time.perf_counter() - t0
            digests["equals_unsharded"] = all(digests.get(k) == v for k, v in ref_d.items())
            if not digests["equals_unsharded"]:
                print(json.dumps({"error": "sharded output differs from the unsharded run", "sharded": digests,
                                  "unsharded": ref_d}), file=sys.stderr)
                sys.stderr.flush()
                os._exit(3)
        dist.barrier()

    if rank != 0:
        dist.destroy_process_group()
        return

    hbm, peak_src = peaks()
    S = 2
    # dominant kernel for the roofline object: the one with the largest device time in the step
    kern = {"bv_scan": st["bv_ms"], "join_count": st["join_ms"], "pair_heavy": st["heavy_ms"], "extract": st["extract_ms"]}
    if st2:
        kern["poa_chain"] = st2["poa_busy_ms"]  # device time with at least one k_poa_chain launch running
    dom = max(kern, key=kern.get)
    bv_alg_bytes = st["bv_pairs"] * (512 * S + 4)  # 512*S+4 bytes per (representative, read) comparison
    bv_gbs = bv_alg_bytes / (st["bv_ms"] * 1e-3) / 1e9 if st["bv_ms"] > 0 else 0.0
    if dom == "poa_chain":
        cells = st2["poa_cells"]
        # Algorithmic bytes per DP cell: the 2-byte traceback code, written once (H/F rows stay in the
        # shared-memory ring).  Kernels of concurrently running units overlap on the device, so the denominator is the
        # device time with at least one POA launch group running (union of the groups' CUDA-event intervals,
        # rtl_stats.poa_busy_ms), not the sum of the overlapping launch durations; `traffic` is what the kernels of
        # THIS step wrote by construction (codes + spilled rows, counted by the library), per launch like `achieved`.
        busy = max(st2["poa_busy_ms"], 1e-6)
        gb = cells * 2 / (busy * 1e-3) / 1e9
        nl = max(1, st2["poa_launches"])
        roof = {"kernel": "k_poa_chain (per-pack CTA: graph update + int16 DP + traceback)", "bound": "hbm", "achieved": gb,
                "peak": hbm,
                "unit": "GB/s", "frac": gb / hbm, "traffic": st2["poa_dram_bytes"] / nl,
                "traffic_source": "counted by the library from the launches of this step (codes + spilled rows); "
                                  "profiles/ holds the ncu dram__bytes of a launch of the same shape",
                "algorithmic_bytes_per_launch": cells * 2 / nl, "peak_source": peak_src,
                "gcups": cells / (busy * 1e-3) / 1e9, "launches": st2["poa_launches"],
                "avg_launch_ms": st2["poa_ms"] / nl, "busy_ms": busy,
                "launch_note": "one launch per CTA width and unit and round; launches of different units overlap, busy_ms is "
                               "the union of their CUDA-event intervals",
                "note": "integer-issue bound, not HBM bound (profiles/): GCUPS against the issue ceiling is the "
                        "meaningful rate; the HBM fraction is reported as SURVEY 8(d) defines it"}
    else:
        roof = {"kernel": "k_bv_scan", "bound": "hbm", "achieved": bv_gbs, "peak": hbm, "unit": "GB/s",
                "frac": bv_gbs / hbm, "traffic": None, "peak_source": peak_src, "pairs": st["bv_pairs"],
                "launches": st["bv_launches"], "avg_launch_ms": st["bv_ms"] / max(1, st["bv_launches"]),
                "note": "algorithmic bytes = pairs x (512*S+4) per SURVEY 8(d); seeds are tiled in shared memory so "
                        "this exceeds DRAM traffic by design (see DESIGN.md)"}
    line = {
        "metric": "reads/sec cluster+correct" if args.correct else "reads/sec cluster",
        "value": n_reads / (ms_res * 1e-3), "unit": "reads/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_res, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": workload_config(args, n_reads, total_genes, world),
        "e2e": {"value": n_reads / (ms_e2e * 1e-3), "unit": "reads/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "note": "same timed loop as `value`; value excludes only the device time of the H2D copy of the read "
                        "set (%.1f ms per step, CUDA events)" % (upload_ms / args.steps)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
        "kernels_ms_per_step": kern,
        "bv_scan": {"pairs": st["bv_pairs"], "ms": st["bv_ms"], "alg_gbs": bv_gbs, "alg_frac_of_hbm": bv_gbs / hbm,
                    "note": "tiled regime (seeds resident in shared memory): algorithmic bytes, not DRAM traffic; the "
                            "streaming regime is measured by tools/bv_stream_bench.py (profiles/)"},
        "library_ms": {"cluster_reads": st["total_ms"], "correct_reads": st2["total_ms"] if st2 else 0.0,
                       "poa_wall": st2["poa_wall_ms"] if st2 else 0.0, "upload": st["upload_ms"]},
        "counters": {"clusters": int(cl.n_clusters), "waves": st["waves"], "rounds": st["rounds"],
                     "full_pairs": st["full_pairs"], "heavy_pairs": st["heavy_pairs"],
                     "poa_cells": st2["poa_cells"] if st2 else 0, "poa_alignments": st2["poa_alignments"] if st2 else 0,
                     **digests},
    }
    # ---- the bitvector scan in its streaming regime (1, 2 and 4 seeds against every read — the regime BASELINE.json's
    # >= 50 %-of-HBM target is about), through the C ABI (rtl_bv_scan).  The read set is the same generator at 400 k reads
    # (410 MB of bitvectors, 3x the L2) and the L2 is flushed before every launch, so every byte comes from HBM.
    if world == 1:
        try:
            big = make_workload(8000)
            ctx.upload(big.bases, big.offsets)
            flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
            targets = np.arange(big.n, dtype=np.int32)
            stream_lines = []
            for ns in (1, 2, 4):
                seeds = np.linspace(0, big.n - 1, ns).astype(np.int32)
                times = []
                for _ in range(7):
                    flush.zero_()
                    torch.cuda.synchronize()
                    ctx.bv_scan(seeds, targets, 0.4, kmer_size=10, is_rna=False, want_output=False)
                    times.append(ctx.stats()["bv_ms"])
                ms = float(np.median(times[2:]))  # (the first launches also extract the k-mers of the new read set)
                gbs = big.n * (512 * S + 4) / (ms * 1e-3) / 1e9  # every read's bitvectors streamed once
                stream_lines.append({"seeds": ns, "kernel_ms": ms, "streamed_GBps": gbs, "frac_of_hbm": gbs / hbm})
            line["bv_scan"]["streaming"] = {"reads": int(big.n), "bytes_per_read": 512 * S + 4,
                                           "l2": "flushed (512 MB memset) before every launch; 410 MB streamed per launch",
                                           "runs": stream_lines}
            del flush
        except Exception as e:
            line["bv_scan"]["streaming"] = {"error": str(e)}
    # ---- CPU baseline on a bounded sample (rank 0, N=1)
    if world == 1 and not args.no_cpu_baseline:
        try:
            import oracle
            cores = os.cpu_count() or 1
            genes = ref_sample_genes(args, steps=1)
            srs = make_workload(genes)
            if oracle.have_ref():
                lib, kind = oracle.reference(), "reference"
            else:
                lib, kind = oracle.oracle(), "port"
            t0 = time.perf_counter()
            c = lib.cluster_reads(srs.bases, srs.offsets, is_rna=False, n_threads=cores)
            if args.correct and kind == "reference":
                lib.correct_reads(srs.bases, srs.quals, srs.offsets, c, n_threads=cores, **CORRECT_KW)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": srs.n / dt, "unit": "reads/s", "cores": cores, "kind": kind,
                                    "sample": "same generator at %d genes x 50 = %d reads, cluster%s, %d threads, %.1f s "
                                              "(clustering cost grows ~quadratically: the full-size rate is lower)" % (
                                                  genes, srs.n, "+correct" if args.correct else "", cores, dt),
                                    "full_config_one_off": full_config_reference()}
        except Exception as e:  # the baseline is reporting only
            line["cpu_baseline"] = {"value": None, "unit": "reads/s", "cores": 0, "kind": "port", "sample": "failed: %s" % e}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
