"""Generates tests/golden/stralg_golden.npz from the UNMODIFIED reference.

Run in the build container only (needs /root/reference and oracle/_ref/libstralg_ref.so, built by
`make -C oracle`):   python tests/golden/make_golden.py

Every output array below is produced by the reference's own functions through ctypes:
sa_is_construction / skew_sa_construction / qsort_sa_construction (suffix_array.h:22-41),
compute_lcp (suffix_array.c:64-85), build_complete_table (bwt.c:134-161) and the exact-match
iterator (bwt.c:164-217).  Inputs are the strings of the reference's own tests
(tests/stralg/match_test.c:682-696, suffix_array_test.c:19-32, bwt_test.c:16-38,
remap_test.c:10-14, tests/stralg/test-data/{modest-proposal,repetitive-string}.txt) plus a few
seeded synthetic texts.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from _oracle import Ref  # noqa: E402

REF_TESTS = "/root/reference/tests/stralg"
STRINGS = ["acacacg", "gacacacag", "acacacag", "acagcaca", "acatgaca", "acgc", "ccgc", "aaaaaaaaa"]
PATTERNS = ["aca", "ac", "ca", "a", "c", "acg", "cg", "g", "cgc", "acgc", "aaa", "aaccaac"]


def table_case(ref, out, name, raw: bytes, patterns):
    t = ref.tables(raw)
    codes = t["codes"]
    sa, isa, lcp = ref.sa_lcp(codes, t["sigma"])
    assert (sa == t["sa"]).all()
    # the reference's own cross-constructor check (match_test.c:479,517,539)
    assert (ref.sa(codes, t["sigma"], "skew") == sa).all()
    assert (ref.sa(codes, t["sigma"], "sa_is_mem") == sa).all()
    out[f"{name}/raw"] = np.frombuffer(raw, dtype=np.uint8)
    out[f"{name}/codes"] = codes
    out[f"{name}/sigma"] = np.array([t["sigma"]], dtype=np.uint32)
    out[f"{name}/sa"] = sa
    out[f"{name}/isa"] = isa
    out[f"{name}/lcp"] = lcp
    out[f"{name}/c"] = t["c"]
    if len(raw) <= 4096:
        out[f"{name}/o"] = t["o"]
    table = t["table"]
    for k, p in enumerate(patterns):
        pc = table[np.frombuffer(p, dtype=np.uint8)]
        if (pc < 0).any():
            continue  # remap() == NULL: the reference skips the pattern (match_test.c:635)
        L, R, pos = ref.exact_matches(t["handle"], pc.astype(np.uint8))
        out[f"{name}/pat{k}/codes"] = pc.astype(np.uint8)
        out[f"{name}/pat{k}/LR"] = np.array([L, R], dtype=np.uint32)
        out[f"{name}/pat{k}/pos"] = pos
    ref.free_tables(t["handle"])


def main():
    ref = Ref()
    out = {}
    for i, s in enumerate(STRINGS):
        table_case(ref, out, f"grid{i}", s.encode(), [p.encode() for p in PATTERNS])
    table_case(ref, out, "mississippi", b"mississippi", [b"ssi", b"i", b"mississippi", b"pp", b"sip"])
    table_case(ref, out, "ababacabac", b"ababacabac", [b"aba", b"c", b"bac"])
    table_case(ref, out, "acagtgtaac", b"acagtgtaac", [b"gt", b"aa"])
    with open(os.path.join(REF_TESTS, "test-data/modest-proposal.txt"), "rb") as f:
        modest = f.read().split(b"\0")[0]
    table_case(ref, out, "modest", modest, [b"the", b"children", b"zzz", b"a"])
    with open(os.path.join(REF_TESTS, "test-data/repetitive-string.txt"), "rb") as f:
        rep = f.read().split(b"\0")[0]
    table_case(ref, out, "repetitive", rep, [b"ababaaba", b"ab", b"b"])

    # seeded synthetic code strings straight into the constructors (no remap: covers sigma = 256)
    rng = np.random.default_rng(20261017)
    synth = {
        "rand_dna_5000": (rng.integers(1, 5, 5000), 5),
        "rand_bin_3000": (rng.integers(1, 3, 3000), 3),
        "rand_byte_4000": (rng.integers(1, 256, 4000), 256),
        "unary_2000": (np.ones(2000, dtype=np.int64), 2),
        "period4_4001": (np.tile(np.array([1, 2, 3, 4]), 1001)[:4001], 5),
        "period37_3000": (np.tile(rng.integers(1, 5, 37), 100)[:3000], 5),
        "fib_2584": (None, 3),
        "single_1": (np.array([3]), 5),
        "empty_0": (np.array([], dtype=np.int64), 5),
    }
    a, b = [1], [1, 2]
    while len(b) < 2584:
        a, b = b, b + a
    synth["fib_2584"] = (np.array(b[:2584]), 3)
    for name, (sym, sigma) in synth.items():
        codes = np.concatenate([sym.astype(np.uint8), np.zeros(1, np.uint8)])
        if len(sym) == 0:
            sa = np.array([0], dtype=np.uint32)
            isa = np.array([0], dtype=np.uint32)
            lcp = np.array([0], dtype=np.uint32)
        else:
            sa, isa, lcp = ref.sa_lcp(codes, sigma)
            if len(sym) > 1:
                assert (ref.sa(codes, sigma, "skew") == sa).all()
                assert (ref.sa(codes, sigma, "sa_is_mem") == sa).all()
        out[f"{name}/codes"] = codes
        out[f"{name}/sigma"] = np.array([sigma], dtype=np.uint32)
        out[f"{name}/sa"] = sa
        out[f"{name}/isa"] = isa
        out[f"{name}/lcp"] = lcp
    np.savez_compressed(os.path.join(HERE, "stralg_golden.npz"), **out)

    # the known-answer vectors the reference's tests spell out literally
    kat = {
        "mississippi": {  # tests/stralg/bwt_test.c:16-38
            "remapped": [2, 1, 4, 4, 1, 4, 4, 1, 3, 3, 1, 0],
            "c_table": [0, 1, 5, 6, 8],
            "o_rows_by_symbol": [
                [0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1],
                [0, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 3, 4],
                [0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1],
                [0, 0, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2],
                [0, 0, 0, 1, 2, 2, 2, 2, 2, 3, 4, 4, 4]],
        },
        "ababacabac": {  # tests/stralg/suffix_array_test.c:19-32
            "sa": [10, 0, 6, 2, 8, 4, 1, 7, 3, 9, 5],
        },
        "acagtgtaac": {  # tests/stralg/remap_test.c:12-14
            "remapped": [1, 2, 1, 3, 4, 3, 4, 1, 1, 2, 0],
        },
    }
    assert out["mississippi/codes"].tolist() == kat["mississippi"]["remapped"]
    assert out["mississippi/c"].tolist() == kat["mississippi"]["c_table"]
    assert out["mississippi/o"].T.tolist() == kat["mississippi"]["o_rows_by_symbol"]
    assert out["ababacabac/sa"].tolist() == kat["ababacabac"]["sa"]
    assert out["acagtgtaac/codes"].tolist() == kat["acagtgtaac"]["remapped"]
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(kat, f, indent=1)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
