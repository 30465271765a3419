#!/usr/bin/env python
"""bench.py — reads/s of RATTLE's cluster(+correct) hot path on synthetic cDNA reads (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--genes G]

A "step" is one full pass of the hot path over the workload: k-mer extraction, the greedy bitvector/k-mer
clustering (initial pass + merge rounds) and POA correction of the resulting clusters (--no-correct: clustering
only).  `value` is measured with the reads already resident in HBM (rtl_reads_upload done before the timed region);
`e2e` goes through the reference-shaped C-ABI call with HOST buffers (H2D of the reads and D2H of the cluster set /
FASTQ text inside the timed region).  N>1 (torchrun) is WEAK scaling: N GPUs cluster and correct N x 100 k reads;
the (seed,target) pair evaluation of every greedy wave is sharded over ranks with one NCCL min-allreduce of the
decision arrays per wave phase, and clusters are sharded over ranks for correction (no collective).

--impl reference times the UNMODIFIED reference (oracle/_ref/libref_shim.so, compiled from /root/reference) on the
host cores, on a bounded sample of the same workload.
"""
import argparse
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
# dram__bytes_read+write of one k_poa_strip launch (600 alignments, ncu --set full, profiles/ncu_poa_strip_r01.txt)
POA_TRAFFIC_PER_LAUNCH = 10.09e9
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
    genes = args.ref_genes
    rs = make_workload(genes)
    sample = "config-2 shape at %d genes x 50 reads = %d reads (~1.5 kb cDNA), cluster%s with %d threads" % (
        genes, rs.n, "+correct" if args.correct else "", cores)
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
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64/int32",
            "data": "synthetic", "config": workload_config(args, rs.n, genes),
            "cpu_baseline": {"value": v, "unit": "reads/s", "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(args, n_reads, genes):
    return {"workload": "BASELINE.json configs[1] per GPU: %d synthetic cDNA reads x ~1.5 kb (%d genes x 50 reads, 3/2/2%% "
                        "sub/ins/del, random strand), k=10 gene clustering%s" % (n_reads, genes,
                                                                               " + correct" if args.correct else ""),
            "n_reads": int(n_reads), "kmer_size": 10, "strands": 2, "l2": "inputs larger than L2 (k-mer lists + "
            "bitvectors of the workload exceed 126 MB); no explicit flush", "parallelism": "pair-shard x%d" % args.gpus}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genes", type=int, default=2000,
                    help="genes PER GPU: 2000 genes x 50 reads = 100 k reads (configs[1]) at N=1; N GPUs cluster and "
                         "correct N x 100 k reads (weak scaling: 8 GPUs = 800 k reads, the size BASELINE.json's metric names)")
    ap.add_argument("--ref-genes", type=int, default=400, help="size of the bounded CPU sample (x50 reads)")
    ap.add_argument("--no-correct", dest="correct", action="store_false", default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="library tunable key=value (rtl_set_option)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if args.correct is None:
            args.correct = True
        reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import rattle_b200

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
    if args.correct is None:
        args.correct = True

    if world > 1:
        from rattle_b200.dist import make_allreduce_callback
        ctx.set_shard(rank, world, make_allreduce_callback(stream.cuda_stream))

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

    def correct_shard(cl):
        """clusters sharded round-robin over ranks (independent packs, no collective)"""
        if not args.correct:
            return None
        if world == 1:
            sub = cl
        else:
            from rattle_b200.dist import shard_clusters
            sub, _ = shard_clusters(cl, rank, world)
        # results stay in the Context's host buffers (what the C ABI wrote): no copy into Python bytes objects
        return ctx.correct_reads(bases_np, quals_np, rs.offsets, sub, as_bytes=False, **CORRECT_KW)

    def step_resident():
        cl = ctx.cluster_resident(**CLUSTER_KW)
        st = ctx.stats()
        out = correct_shard(cl)
        st2 = ctx.stats() if args.correct else None
        return cl, st, st2, out

    def step_e2e():
        cl = ctx.cluster_reads(bases_np, rs.offsets, **CLUSTER_KW)
        st = ctx.stats()
        out = correct_shard(cl)
        return cl, st, out

    def timed(fn, steps, warmup, sampler=None):
        res = None
        for _ in range(warmup):
            res = fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stats = []
        if sampler:
            sampler.__enter__()
        e0.record(stream)
        for _ in range(steps):
            res = fn()
            stats.append(res)
        e1.record(stream)
        barrier()
        if sampler:
            sampler.__exit__()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, stats

    # ---- device-resident arm (`value`)
    ctx.upload(bases_np, rs.offsets)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_res, res_stats = timed(step_resident, args.steps, args.warmup, sampler)
    clocks = sampler.summary() if sampler else None
    # ---- end-to-end arm through the C ABI with host buffers
    ms_e2e, e2e_stats = timed(step_e2e, args.steps, 1)

    cl, st, st2, out = res_stats[-1]
    e_cl, e_st, e_out = e2e_stats[-1]
    launches = st["kernel_launches"] + (st2["kernel_launches"] if st2 else 0)
    h2d = e_st["h2d_bytes"]
    d2h = e_st["d2h_bytes"] + int(cl.n_clusters) * 13 + n_reads * 5
    if args.correct and e_out is not None:
        h2d += 2 * int(rs.offsets[-1])
        d2h += sum(len(x) for x in e_out)

    if rank != 0:
        dist.destroy_process_group()
        return

    hbm, peak_src = peaks()
    S = 2
    # dominant kernel for the roofline object: the one with the largest device time in the step
    kern = {"bv_scan": st["bv_ms"], "join_count": st["join_ms"], "pair_heavy": st["heavy_ms"], "extract": st["extract_ms"]}
    if st2:
        kern["poa"] = st2["poa_ms"]
    dom = max(kern, key=kern.get)
    roof = None
    bv_alg_bytes = st["bv_pairs"] * (512 * S + 4)  # 512*S+4 bytes per (representative, read) comparison
    bv_gbs = bv_alg_bytes / (st["bv_ms"] * 1e-3) / 1e9 if st["bv_ms"] > 0 else 0.0
    if dom == "poa":
        cells = st2["poa_cells"]
        # Algorithmic bytes per DP cell: the 2-byte traceback code, written once (H/F rows stay in the
        # shared-memory ring).  Units (concurrent launch groups) overlap on the device, so the denominator is the
        # device time with at least one POA launch group running (union of the groups' CUDA-event intervals,
        # rtl_stats.poa_busy_ms), not the sum of the overlapping launch durations.
        busy = max(st2["poa_busy_ms"], 1e-6)
        gb = cells * 2 / (busy * 1e-3) / 1e9
        roof = {"kernel": "k_poa_strip (+ k_poa_strip_traceback)", "bound": "hbm", "achieved": gb, "peak": hbm,
                "unit": "GB/s", "frac": gb / hbm, "traffic": POA_TRAFFIC_PER_LAUNCH, "peak_source": peak_src,
                "gcups": cells / (busy * 1e-3) / 1e9, "launches": st2["poa_launches"],
                "avg_launch_ms": st2["poa_ms"] / max(1, st2["poa_launches"]), "busy_ms": busy,
                # from the committed ncu --set full capture of one 600-alignment launch (profiles/ncu_poa_strip_r01.txt)
                "ncu": {"issue_slot_utilisation": 0.757, "dram_throughput_pct": 13.5, "warps_per_sm": 24,
                        "registers": 72, "dram_bytes_per_cell": 2.0, "launch_algorithmic_bytes": 1.00e10,
                        "launch_dram_bytes": 1.009e10, "launch_ms": 9.12},
                "note": "integer-issue bound, not HBM bound: ~50 SASS instructions per DP cell at ~75 % issue-slot "
                        "utilisation (profiles/); GCUPS is the meaningful rate"}
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
        "vs_baseline": None, "dtype": "u64/int32", "data": "synthetic",
        "config": workload_config(args, n_reads, total_genes),
        "e2e": {"value": n_reads / (ms_e2e * 1e-3), "unit": "reads/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(launches) * args.steps,
        "clocks": clocks,
        "roofline": roof,
        "kernels_ms_per_step": kern,
        "bv_scan": {"pairs": st["bv_pairs"], "ms": st["bv_ms"], "alg_gbs": bv_gbs, "alg_frac_of_hbm": bv_gbs / hbm},
        "library_ms": {"cluster_reads": st["total_ms"], "correct_reads": st2["total_ms"] if st2 else 0.0,
                       "poa_wall": st2["poa_wall_ms"] if st2 else 0.0},
        "counters": {"clusters": int(cl.n_clusters), "waves": st["waves"], "rounds": st["rounds"],
                     "full_pairs": st["full_pairs"], "heavy_pairs": st["heavy_pairs"],
                     "poa_cells": st2["poa_cells"] if st2 else 0, "poa_alignments": st2["poa_alignments"] if st2 else 0},
    }
    # ---- CPU baseline on a bounded sample (rank 0, N=1)
    if world == 1 and not args.no_cpu_baseline:
        try:
            import oracle
            cores = os.cpu_count() or 1
            srs = make_workload(args.ref_genes)
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
                                              "(clustering cost grows ~quadratically: the 100 k-read rate is lower)" % (
                                                  args.ref_genes, srs.n, "+correct" if args.correct else "", cores, dt)}
        except Exception as e:  # the baseline is reporting only
            line["cpu_baseline"] = {"value": None, "unit": "reads/s", "cores": 0, "kind": "port", "sample": "failed: %s" % e}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
