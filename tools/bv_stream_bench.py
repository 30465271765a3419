"""Bitvector-scan kernel (k_bv_scan, cluster.cpp:13-19,43) in its two regimes, through the C ABI entry `rtl_bv_scan`:

  streaming   1-4 representatives against every read: each read's 512*S-byte bitvector is streamed from HBM once and
              the kernel is bandwidth bound — the regime the >= 50 %-of-HBM target of BASELINE.json is about;
  tiled       128+ representatives resident in shared memory per streamed read: POPC/issue bound, reported as
              pairs/s and as algorithmic GB/s (pairs x (512*S+4) B), which exceeds DRAM traffic by design.

    python tools/bv_stream_bench.py [--genes 8000] [--seeds 1,2,4,8,32,128,512] [--rna]

The read set must be larger than L2 (126 MB): 8000 genes x 50 reads = 400 k reads = 205 MB of bitvectors per strand.
Prints one JSON line per seed count; kernel time = CUDA events around the launch (rtl_stats.bv_ms).
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tools import synth  # noqa: E402


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genes", type=int, default=8000)
    ap.add_argument("--seeds", default="1,2,4,8,32,128,512")
    ap.add_argument("--rna", action="store_true")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--kernel", type=int, default=0, help="0 = k_bv_scan (default path), 2 = bulk-copy ring scan k_bv_stream for at most 16 seeds")
    args = ap.parse_args()
    import rattle_b200
    rs = synth.config2(n_genes=args.genes).sorted_by_length()[0]
    S = 1 if args.rna else 2
    ctx = rattle_b200.Context(0)
    ctx.set_option("bv_kernel", args.kernel)
    tile = 128
    ctx.upload(rs.bases, rs.offsets)
    targets = np.arange(rs.n, dtype=np.int32)
    hbm, src = peak()
    for ns in [int(x) for x in args.seeds.split(",")]:
        seeds = np.linspace(0, rs.n - 1, ns).astype(np.int32)
        best = None
        for _ in range(args.reps):
            ctx.bv_scan(seeds, targets, 0.4, kmer_size=10, is_rna=args.rna, want_output=False)
            st = ctx.stats()
            best = st["bv_ms"] if best is None else min(best, st["bv_ms"])
        pairs = ns * rs.n
        tiles = (ns + tile - 1) // tile
        dram = tiles * rs.n * (512 * S + 4)  # every seed tile streams every read's bitvectors once
        print(json.dumps({"kernel": "k_bv_stream" if (args.kernel == 2 and ns <= 16) else "k_bv_scan", "seeds": ns, "reads": rs.n, "strands": S, "kernel_ms": best,
                          "pairs_per_s": pairs / (best * 1e-3),
                          "streamed_GBps": dram / (best * 1e-3) / 1e9, "streamed_frac_of_hbm": dram / (best * 1e-3) / 1e9 / hbm,
                          "algorithmic_GBps": pairs * (512 * S + 4) / (best * 1e-3) / 1e9, "hbm_peak": hbm, "peak_source": src}))


if __name__ == "__main__":
    main()
