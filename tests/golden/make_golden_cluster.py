"""Small function-level golden vectors for hot path A from the UNMODIFIED reference (oracle/_ref/libref_shim.so).
Run in the build container: python tests/golden/make_golden_cluster.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import oracle  # noqa: E402
from tools import synth  # noqa: E402

ref = oracle.reference()
HERE = os.path.dirname(os.path.abspath(__file__))
out = {"pairs": [], "cluster": []}
rs = synth.generate(seed=31, n_genes=4, reads_per_tx=4, len_mean=260.0, len_sd=40.0, len_min=150, len_max=400)
rs = rs.sorted_by_length()[0]
seqs = [rs.seq(i).decode() for i in range(rs.n)]
for k, t_s, t_v, thr, rna in [(10, 0.2, 1e6, 0.4, 0), (11, 0.3, 25.0, 0.3, 0), (6, 0.5, 25.0, 0.4, 1), (10, 0.2, 1e6, 0.0, 0)]:
    res = []
    for i in range(rs.n):
        for j in range(i + 1, rs.n):
            res.append(ref.pair_match(rs.seq(i), rs.seq(j), k, t_s, t_v, thr, rna))
    out["pairs"].append({"k": k, "t_s": t_s, "t_v": t_v, "thr": thr, "is_rna": rna, "match": res})
out["seqs"] = seqs
# k-mer lists / bitvectors of the first three reads
km = []
for i in range(3):
    n, fh, fp, rh, rp, bf, br = ref.extract_kmers(rs.seq(i), 10, True)
    km.append({"fh": fh.tolist(), "fp": fp.tolist(), "rh": rh.tolist(), "rp": rp.tolist(), "bf": [int(x) for x in bf],
               "br": [int(x) for x in br]})
out["kmers_k10"] = km
for seed, genes, per, rna, kw in [(41, 12, 10, False, {}), (42, 10, 8, True, {}),
                                  (43, 6, 10, False, dict(k=11, t_s=0.3, t_v=25.0)),
                                  (44, 8, 6, False, dict(k=6, t_s=0.5, t_v=25.0, bv_thr=0.4, bv_min=0.4))]:
    r = synth.generate(seed=seed, n_genes=genes, n_isoforms=2 if seed == 43 else 1, reads_per_tx=per, len_mean=700.0,
                       len_sd=120.0, len_min=300, len_max=1500).sorted_by_length()[0]
    cl = ref.cluster_reads(r.bases, r.offsets, is_rna=rna, n_threads=4, **kw)
    out["cluster"].append({"synth": dict(seed=seed, n_genes=genes, n_isoforms=2 if seed == 43 else 1, reads_per_tx=per,
                                         len_mean=700.0, len_sd=120.0, len_min=300, len_max=1500),
                           "is_rna": rna, "kw": kw, "n_clusters": int(cl["n_clusters"]),
                           "main_id": cl["main_id"].tolist(), "main_rev": cl["main_rev"].tolist(),
                           "cl_off": cl["cl_off"].tolist(), "mem_id": cl["mem_id"].tolist(),
                           "mem_rev": cl["mem_rev"].tolist()})
json.dump(out, open(os.path.join(HERE, "cluster_small.json"), "w"))
print("ok", len(out["pairs"][0]["match"]), [c["n_clusters"] for c in out["cluster"]])
