"""Golden digests of the UNMODIFIED reference CLI (oracle/_ref/rattle, compiled from /root/reference by
oracle/Makefile) on real Nanopore reads: the first 1500 records of the reference's own toy data set
(/root/reference/toyset/rna/input/sample.fastq), committed as tests/golden/toyset_rna_1500.fastq.gz.

    python tests/golden/make_golden_cli.py          (needs /root/reference and oracle/_ref/rattle)

The drop-in CLI (integration/_build/rattle = reference main.cpp/fasta.cpp/utils.cpp + librattle_b200) must reproduce
every file byte for byte (corrected.fq as a multiset of records: the reference's order depends on -t).
"""
import gzip
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
FIX = os.path.join(HERE, "toyset_rna_1500.fastq.gz")
REF_FASTQ = "/root/reference/toyset/rna/input/sample.fastq"

# (name, subcommand arguments after `-i <input> -o <dir>`, files to digest)
STEPS = [
    ("cluster_rna", ["cluster", "--rna", "-t", "8"], ["clusters.out"]),
    ("correct", ["correct", "-t", "1"], ["consensi.fq", "uncorrected.fq", "corrected.fq"]),
    ("polish_rna", ["polish", "--rna", "-t", "8"], ["transcriptome.fq"]),
    ("cluster_rna_iso", ["cluster", "--rna", "--iso", "-t", "8"], ["clusters.out"]),
    ("cluster_cdna", ["cluster", "-t", "8"], ["clusters.out"]),
]


def sha(path, as_multiset=False):
    data = open(path, "rb").read()
    if as_multiset:
        lines = data.split(b"\n")
        recs = sorted(b"\n".join(lines[i:i + 4]) for i in range(0, len(lines) - 1, 4))
        data = b"\n".join(recs)
    return hashlib.sha256(data).hexdigest()


def run_pipeline(binary, fastq, workdir):
    """runs STEPS with `binary`; returns {step: {file: sha256}}"""
    out = {}
    clusters_rna = None
    for name, argv, files in STEPS:
        d = os.path.join(workdir, name)
        os.makedirs(d, exist_ok=True)
        if argv[0] == "cluster":
            cmd = [binary, "cluster", "-i", fastq, "-o", d] + argv[1:]
        elif argv[0] == "correct":
            cmd = [binary, "correct", "-i", fastq, "-c", clusters_rna, "-o", d] + argv[1:]
        else:
            cmd = [binary, "polish", "-i", os.path.join(workdir, "correct", "consensi.fq"), "-o", d] + argv[1:]
        subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        if name == "cluster_rna":
            clusters_rna = os.path.join(d, "clusters.out")
        out[name] = {f: sha(os.path.join(d, f), as_multiset=(f == "corrected.fq")) for f in files}
    return out


def unpack_fixture(workdir):
    fastq = os.path.join(workdir, "toy.fastq")
    with gzip.open(FIX, "rb") as src, open(fastq, "wb") as dst:
        shutil.copyfileobj(src, dst)
    return fastq


if __name__ == "__main__":
    if not os.path.exists(FIX):
        lines = []
        with open(REF_FASTQ, "rb") as f:
            for _ in range(6000):
                lines.append(f.readline())
        with gzip.GzipFile(FIX, "wb", compresslevel=9, mtime=0) as g:
            g.write(b"".join(lines))
    ref = os.path.join(ROOT, "oracle", "_ref", "rattle")
    with tempfile.TemporaryDirectory() as wd:
        res = run_pipeline(ref, unpack_fixture(wd), wd)
    json.dump({"fixture": os.path.basename(FIX), "reads": 1500, "reference_commit": "a892888", "digests": res},
              open(os.path.join(HERE, "cli_toyset.json"), "w"), indent=1)
    print(json.dumps(res, indent=1))
