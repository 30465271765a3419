"""Golden digests from the UNMODIFIED reference (oracle/_ref, compiled from /root/reference by oracle/Makefile) for the
larger parity cases.  Run in the build container:

    python tests/golden/make_golden_big.py [toyset] [config2] [config4] [hps]

  toyset  : the reference's whole toy set (/root/reference/toyset/rna/input/sample.fastq, 8306 records; committed
            as tests/golden/toyset_rna_full.fastq.gz) through the reference CLI: cluster --rna, correct -t 1, polish,
            cluster --rna --iso, cluster (cDNA).  -> cli_toyset_full.json (sha256 + md5; SURVEY.md 8(c) quotes the md5s)
  config2 : cluster + correct on the bench generator at 400 genes x 50 reads (20 k reads), through the shim
            (cluster.cpp:93, correct.cpp:311).  Every cluster has <= split reads, so consensi/uncorrected do not depend
            on the thread count; corrected.fq is compared as a multiset.  -> config2_400_correct.json
  config4 : correct on 200 clusters x 32 forward reads x 2 kb (BASELINE.json configs[3] shape) -> config4_200.json
  hps     : reference-written clusters.out (gene level and --iso) of the 1500-read fixture, committed as binary
            fixtures for the codec test -> clusters_rna_1500.out, clusters_rna_iso_1500.out
"""
import gzip
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import numpy as np  # noqa: E402

REF_FASTQ = "/root/reference/toyset/rna/input/sample.fastq"
FULL_FIX = os.path.join(HERE, "toyset_rna_full.fastq.gz")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "rattle")


def fq_digests(out):
    """(corrected, uncorrected, consensi) bytes -> digests; corrected as a sorted multiset of 4-line records"""
    def multiset(data):
        lines = data.split(b"\n")
        recs = sorted(b"\n".join(lines[i:i + 4]) for i in range(0, len(lines) - 1, 4))
        return b"\n".join(recs)
    return {"corrected_sorted": hashlib.sha256(multiset(out[0])).hexdigest(),
            "uncorrected": hashlib.sha256(out[1]).hexdigest(),
            "consensi": hashlib.sha256(out[2]).hexdigest(),
            "bytes": [len(out[0]), len(out[1]), len(out[2])]}


def config4_set(n_clusters, reads_per=32, length=2000):
    """cluster c = reads 32c..32c+31 in length-descending order, rev=false, gene_id=-1"""
    from tools import synth
    rs = synth.generate(seed=42, n_genes=n_clusters, n_isoforms=1, reads_per_tx=reads_per, len_mean=float(length),
                        len_sd=0.0, len_min=length, len_max=length, p_flip=0.0, shuffle=False)
    lens = rs.lengths()
    order = []
    for c in range(n_clusters):
        idx = np.arange(c * reads_per, (c + 1) * reads_per)
        order.append(idx[np.argsort(-lens[idx], kind="stable")])
    rs = rs.take(np.concatenate(order))
    n = n_clusters * reads_per
    ids = np.arange(n, dtype=np.int32)
    off = np.arange(0, n + 1, reads_per, dtype=np.int64)
    cl = dict(n_clusters=n_clusters, main_id=ids[off[:-1]].copy(), main_rev=np.zeros(n_clusters, np.uint8), cl_off=off,
              mem_id=ids, mem_rev=np.zeros(n, np.uint8))
    return rs, cl


def do_toyset():
    import make_golden_cli as cli
    if not os.path.exists(FULL_FIX):
        with open(REF_FASTQ, "rb") as f, gzip.GzipFile(FULL_FIX, "wb", compresslevel=9, mtime=0) as g:
            shutil.copyfileobj(f, g)
    with tempfile.TemporaryDirectory() as wd:
        fastq = os.path.join(wd, "toy.fastq")
        with gzip.open(FULL_FIX, "rb") as src, open(fastq, "wb") as dst:
            shutil.copyfileobj(src, dst)
        t0 = time.time()
        res = cli.run_pipeline(REF_BIN, fastq, wd)
        md5 = {}
        for name, _, files in cli.STEPS:
            for f in files:
                data = open(os.path.join(wd, name, f), "rb").read()
                if f == "corrected.fq":
                    data = b"".join(sorted(data.splitlines(keepends=True)))  # `sort corrected.fq` of SURVEY 8(c)
                md5.setdefault(name, {})[f] = hashlib.md5(data).hexdigest()
        dt = time.time() - t0
    out = {"fixture": os.path.basename(FULL_FIX), "reads": 8306, "reference_commit": "a892888", "digests": res,
           "md5": md5, "reference_seconds": dt}
    json.dump(out, open(os.path.join(HERE, "cli_toyset_full.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


def do_config2(genes=400):
    import oracle
    from tools import synth
    rs = synth.config2(n_genes=genes).sorted_by_length()[0]
    ref = oracle.reference()
    t0 = time.time()
    cl = ref.cluster_reads(rs.bases, rs.offsets, is_rna=False, n_threads=os.cpu_count())
    t1 = time.time()
    sizes = np.diff(cl["cl_off"])
    assert sizes.max() <= 200, "a multi-pack cluster would make consensi depend on the thread count"
    out = ref.correct_reads(rs.bases, rs.quals, rs.offsets, cl, n_threads=os.cpu_count(), min_occ=0.3, gap_occ=0.3,
                            err_ratio=30.0, split=200, min_reads=5)
    t2 = time.time()
    res = {"genes": genes, "n_reads": rs.n, "n_clusters": int(cl["n_clusters"]), "digests": fq_digests(out),
           "cluster_seconds": t1 - t0, "correct_seconds": t2 - t1, "threads": os.cpu_count()}
    json.dump(res, open(os.path.join(HERE, "config2_%d_correct.json" % genes), "w"), indent=1)
    print(res)


def do_config4(n_clusters=200):
    import oracle
    rs, cl = config4_set(n_clusters)
    ref = oracle.reference()
    t0 = time.time()
    out = ref.correct_reads(rs.bases, rs.quals, rs.offsets, cl, n_threads=os.cpu_count(), min_occ=0.3, gap_occ=0.3,
                            err_ratio=30.0, split=200, min_reads=5)
    dt = time.time() - t0
    res = {"clusters": n_clusters, "reads_per_cluster": 32, "read_len": 2000, "n_reads": rs.n,
           "digests": fq_digests(out), "correct_seconds": dt, "threads": os.cpu_count(),
           "input_sha256": hashlib.sha256(rs.bases.tobytes() + rs.quals.tobytes() + rs.offsets.tobytes()).hexdigest()}
    json.dump(res, open(os.path.join(HERE, "config4_%d.json" % n_clusters), "w"), indent=1)
    print(res)


def do_hps():
    import make_golden_cli as cli
    with tempfile.TemporaryDirectory() as wd:
        fastq = cli.unpack_fixture(wd)
        for name, extra in (("clusters_rna_1500.out", []), ("clusters_rna_iso_1500.out", ["--iso"])):
            d = os.path.join(wd, name + ".d")
            os.makedirs(d)
            subprocess.run([REF_BIN, "cluster", "-i", fastq, "-o", d, "--rna", "-t", "8"] + extra, check=True,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            shutil.copy(os.path.join(d, "clusters.out"), os.path.join(HERE, name))
            print(name, os.path.getsize(os.path.join(HERE, name)))


if __name__ == "__main__":
    what = sys.argv[1:] or ["hps", "config4", "config2", "toyset"]
    for w in what:
        {"toyset": do_toyset, "config2": do_config2, "config4": do_config4, "hps": do_hps}[w]()
