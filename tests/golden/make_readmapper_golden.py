"""Generates tests/golden/readmapper/ from the UNMODIFIED reference read mapper.

Run in the build container only (needs /root/reference and gcc):
    python tests/golden/make_readmapper_golden.py

The reference's `bwt_readmapper` (tools/readmappers/bwt_readmapper/bwt_readmapper.c) is compiled
from the sources where they lie into a scratch directory, run with `-p` on a seeded synthetic
two-record FASTA and with `-d 0` on a seeded FASTQ (exact matching through its approximate
iterator with zero edits, bwt_readmapper.c:128-161), then with `-d 1` / `-d 2` on the longer reads.  Committed: the two input files, the SAM
output, and the sha256 + size of the `.bwttables` file `-p` wrote (the file itself is 0.5 MB).
"""
import hashlib
import json
import os
import shutil
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "readmapper")
REF = "/root/reference"


def main():
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(20261017)
    # two records; the second one carries a repeat of a piece of the first and a run of A
    a = rng.integers(0, 4, 12000)
    b = np.concatenate([rng.integers(0, 4, 5000), a[3000:3400], np.zeros(40, dtype=np.int64), rng.integers(0, 4, 2500)])
    recs = [("chrA", a), ("chrB", b)]
    letters = np.frombuffer(b"ACGT", dtype=np.uint8)
    fasta = os.path.join(OUT, "ref.fa")
    with open(fasta, "w") as f:
        for name, s in recs:
            f.write(f">{name}\n")
            txt = letters[s].tobytes().decode()
            for i in range(0, len(txt), 60):
                f.write(txt[i:i + 60] + "\n")
    fastq = os.path.join(OUT, "reads.fq")
    with open(fastq, "w") as f:
        for k in range(120):
            m = int(rng.integers(8, 40))
            kind = k % 4
            if kind == 3:
                s = rng.integers(0, 4, m)                       # random: mostly no hit
            else:
                src = recs[k % 2][1]
                st = int(rng.integers(0, len(src) - m))
                s = src[st:st + m].copy()
                if kind == 2:
                    s[int(rng.integers(0, m))] ^= 1             # one substitution
            if k == 7:
                s = np.zeros(12, dtype=np.int64)                # poly-A: many hits in chrB
            seq = letters[s].tobytes().decode()
            if k == 11:
                seq = seq[:5] + "N" + seq[6:]                   # a letter no record has: skipped (remap == NULL)
            f.write(f"@read{k}\n{seq}\n+\n{'~' * len(seq)}\n")
    tmp = tempfile.mkdtemp()
    exe = os.path.join(tmp, "bwt_readmapper")
    srcs = [os.path.join(REF, "stralg", x) for x in os.listdir(os.path.join(REF, "stralg")) if x.endswith(".c")]
    srcs += [os.path.join(REF, "bioinf", x) for x in os.listdir(os.path.join(REF, "bioinf")) if x.endswith(".c")]
    srcs.append(os.path.join(REF, "tools/readmappers/bwt_readmapper/bwt_readmapper.c"))
    subprocess.run(["gcc", "-O2", "-std=c11", "-D_GNU_SOURCE", "-w", f"-I{REF}/stralg", f"-I{REF}/bioinf", *srcs,
                    "-o", exe], check=True)
    work = os.path.join(tmp, "ref.fa")
    shutil.copy(fasta, work)
    subprocess.run([exe, "-p", work], check=True, stderr=subprocess.DEVNULL, stdout=subprocess.DEVNULL)
    sam = subprocess.run([exe, "-d", "0", work, fastq], check=True, stderr=subprocess.DEVNULL,
                         stdout=subprocess.PIPE).stdout
    open(os.path.join(OUT, "expected.sam"), "wb").write(sam)
    # approximate matching (edit distance 1 and 2) on the reads of at least 16 / 24 symbols: the
    # tool's D-table-pruned recursion, every (position, CIGAR) in its report order
    approx = {}
    lines = open(fastq).read().split("\n")
    for d, minlen in ((1, 16), (2, 24)):
        sub = os.path.join(OUT, f"reads_d{d}.fq")
        with open(sub, "w") as f:
            for k in range(0, len(lines) - 3, 4):
                if len(lines[k + 1]) >= minlen:
                    f.write("\n".join(lines[k:k + 4]) + "\n")
        out = subprocess.run([exe, "-d", str(d), work, sub], check=True, stderr=subprocess.DEVNULL,
                             stdout=subprocess.PIPE).stdout
        open(os.path.join(OUT, f"expected_d{d}.sam"), "wb").write(out)
        approx[f"d{d}_sam_lines"] = out.count(b"\n")
    tables = open(work + ".bwttables", "rb").read()
    json.dump({"bwttables_sha256": hashlib.sha256(tables).hexdigest(), "bwttables_bytes": len(tables),
               "sam_lines": sam.count(b"\n"), **approx,
               "command": "bwt_readmapper -p ref.fa; bwt_readmapper -d 0 ref.fa reads.fq; "
                          "bwt_readmapper -d 1 ref.fa reads_d1.fq; bwt_readmapper -d 2 ref.fa reads_d2.fq"},
              open(os.path.join(OUT, "meta.json"), "w"), indent=1)
    print(f"{sam.count(10)} SAM lines, .bwttables {len(tables)} bytes")
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
