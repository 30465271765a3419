"""Generates tests/golden/config2_<genes>.json: digest of the UNMODIFIED reference's cluster_reads output
(oracle/_ref/libref_shim.so -> /root/reference/cluster.cpp:93) on the bench workload (tools/synth.config2).
Run in the build container (needs oracle/_ref): python tests/golden/make_golden_config2.py [genes]"""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import oracle  # noqa: E402
from tools import synth  # noqa: E402


def digest(cl):
    h = hashlib.sha256()
    nc = int(cl["n_clusters"])
    for k in ("main_id", "main_rev", "cl_off", "mem_id", "mem_rev"):
        a = np.ascontiguousarray(cl[k])
        if k in ("main_id", "main_rev"):
            a = a[:nc]
        elif k == "cl_off":
            a = a[:nc + 1]
        h.update(a.tobytes())
    return h.hexdigest()


if __name__ == "__main__":
    genes = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    rs = synth.config2(n_genes=genes).sorted_by_length()[0]
    t0 = time.time()
    cl = oracle.reference().cluster_reads(rs.bases, rs.offsets, is_rna=False, n_threads=os.cpu_count())
    dt = time.time() - t0
    out = {"genes": genes, "n_reads": rs.n, "n_clusters": int(cl["n_clusters"]), "sha256": digest(cl),
           "reference_seconds": dt, "threads": os.cpu_count(),
           "input_sha256": hashlib.sha256(rs.bases.tobytes() + rs.offsets.tobytes()).hexdigest()}
    with open(os.path.join(ROOT, "tests", "golden", "config2_%d.json" % genes), "w") as f:
        json.dump(out, f, indent=1)
    print(out)
