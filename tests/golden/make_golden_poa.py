"""Golden vectors for hot path B from the UNMODIFIED reference (oracle/_ref/libref_shim.so = spoa + correct.cpp
compiled from /root/reference).  Run in the build container: python tests/golden/make_golden_poa.py"""
import hashlib
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

cases = []
for seed, n, length, kw in [(11, 5, 80.0, {}), (12, 8, 150.0, dict(p_sub=0.08, p_ins=0.05, p_del=0.05)), (13, 10, 300.0, {})]:
    rs = synth.generate(seed=seed, n_genes=1, reads_per_tx=n, len_mean=length, len_sd=0.0, len_min=int(length),
                        len_max=int(length), p_flip=0.0, shuffle=False, **kw)
    rows = ref.poa_msa(rs.bases, rs.offsets)
    cases.append({"seqs": [rs.seq(i).decode() for i in range(rs.n)], "msa": [r.decode() for r in rows]})
# the spoa test data the reference ships (spoa/test/data/sample.fastq, 55 reads) with RATTLE's scoring (kSW,5,-4,-8,-6;
# the LocalAffine* cases of spoa_test.cpp:115-134,495-516): first 12 reads
fq = "/root/reference/spoa/test/data/sample.fastq"
if os.path.exists(fq):
    rs = synth.read_fastq(fq)
    seqs = [rs.seq(i) for i in range(12)]
    r2 = synth.from_sequences(seqs)
    rows = ref.poa_msa(r2.bases, r2.offsets)
    cases.append({"seqs": [s.decode() for s in seqs], "msa": [r.decode() for r in rows]})
json.dump(cases, open(os.path.join(HERE, "poa_msa.json"), "w"))

sp = dict(seed=21, n_genes=5, reads_per_tx=10, len_mean=350.0, len_sd=0.0, len_min=350, len_max=350, p_flip=0.0,
          shuffle=False)
rs = synth.generate(**sp)
sizes = [10, 10, 10, 10, 6, 4]
off = np.zeros(len(sizes) + 1, np.int64)
off[1:] = np.cumsum(sizes)
ids = np.arange(int(off[-1]), dtype=np.int32)
cl = dict(n_clusters=len(sizes), main_id=ids[off[:-1]].copy(), main_rev=np.zeros(len(sizes), np.uint8), cl_off=off,
          mem_id=ids, mem_rev=np.zeros(len(ids), np.uint8))
kw = dict(min_occ=0.3, gap_occ=0.3, err_ratio=30.0, split=200, min_reads=5)
out = ref.correct_reads(rs.bases, rs.quals, rs.offsets, cl, n_threads=1, **kw)
json.dump({"synth": sp, "sizes": sizes, "kw": kw, "corrected_sha256": hashlib.sha256(out[0]).hexdigest(),
           "uncorrected_sha256": hashlib.sha256(out[1]).hexdigest(), "consensi_sha256": hashlib.sha256(out[2]).hexdigest(),
           "consensi_head": out[2].decode().splitlines()[:2]}, open(os.path.join(HERE, "correct_small.json"), "w"), indent=1)
print("ok", len(cases), len(out[2]))
