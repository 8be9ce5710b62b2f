"""Generates tests/golden/approx_golden.json from the UNMODIFIED reference (needs /root/reference and
oracle/_ref/libstralg_ref.so):   python tests/golden/make_approx_golden.py

For the reference's own test strings and patterns (tests/stralg/match_test.c:682-696 plus
"mississippi", tests/stralg/bwt_test.c:16) and edit distances 0, 1, 2: what init_bwt_approx_iter
collects (interval list, matched lengths, CIGARs, in report order; bwt.c:302-382) and every
(position, CIGAR, matched length) next_bwt_approx_match yields (bwt.c:384-401), on the table
build_complete_table(text, true) makes (with the reverse O table, so the D table prunes).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from _oracle import Ref  # noqa: E402

STRINGS = ["acacacg", "gacacacag", "acacacag", "acagcaca", "acatgaca", "acgc", "ccgc", "aaaaaaaaa", "mississippi"]
PATTERNS = ["aca", "ac", "ca", "a", "c", "acg", "cg", "g", "cgc", "acgc", "aaa", "aaccaac", "ssi", "is", "ippi"]


def main():
    ref = Ref()
    out = []
    for text in STRINGS:
        t = ref.tables(text.encode(), include_reverse=True)
        table = t["table"]
        for pat in PATTERNS:
            pc = table[np.frombuffer(pat.encode(), dtype=np.uint8)]
            if (pc <= 0).any():
                continue  # remap() == NULL: the reference's callers skip the pattern
            for d in (0, 1, 2):
                L, R, ml, cig, hits = ref.approx_matches(t["handle"], pc.astype(np.uint8), d)
                out.append({"text": text, "pattern": pat, "edits": d, "codes": [int(x) for x in pc],
                            "L": [int(x) for x in L], "R": [int(x) for x in R], "match_length": [int(x) for x in ml],
                            "cigars": cig, "hits": [[p, c, m] for p, c, m in hits]})
        ref.free_tables(t["handle"])
    path = os.path.join(HERE, "approx_golden.json")
    json.dump(out, open(path, "w"), separators=(",", ":"))
    print(f"{len(out)} cases, {sum(len(c['L']) for c in out)} intervals, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
