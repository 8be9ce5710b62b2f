"""Packs the reference's 500 kbp human-genome sample (tools/readmappers/data/genomes/hg38-10000.fa: one
record, upper-case ACGT) into 2 bits per base -> tests/golden/hg38_10000.2bit.npy.  The file is DATA
(a public genome excerpt), used by bench.py / the GPU tests to synthesise genome-like texts with real
repeat structure; /root/reference does not exist on the GPU box, so the packed copy is committed."""
import os
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/tools/readmappers/data/genomes/hg38-10000.fa"
HERE = os.path.dirname(os.path.abspath(__file__))

seq = b"".join(ln.strip() for ln in open(REF, "rb") if not ln.startswith(b">"))
a = np.frombuffer(seq, dtype=np.uint8)
lut = np.full(256, 255, dtype=np.uint8)
for k, ch in enumerate(b"ACGT"):
    lut[ch] = k
sym = lut[a]
assert (sym < 4).all(), "non-ACGT letter in the sample"
n = len(sym)
pad = (-n) % 4
s4 = np.concatenate([sym, np.zeros(pad, np.uint8)]).reshape(-1, 4)
packed = (s4[:, 0] | (s4[:, 1] << 2) | (s4[:, 2] << 4) | (s4[:, 3] << 6)).astype(np.uint8)
np.save(os.path.join(HERE, "hg38_10000.2bit.npy"), np.concatenate([np.frombuffer(np.uint32(n).tobytes(), np.uint8), packed]))
print(n, "bases ->", len(packed) + 4, "bytes")
