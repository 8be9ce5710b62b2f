"""CPU tests: the restatement oracle against the reference's golden vectors and, when the
prebuilt reference library is present, against the reference itself."""
import json
import os

import numpy as np
import pytest

from conftest import ROOT, golden_cases


def test_kat_literals(oracle):
    """Known answers spelled out in the reference's tests (bwt_test.c:16-38,
    suffix_array_test.c:19-32, remap_test.c:12-14)."""
    kat = json.load(open(os.path.join(ROOT, "tests", "golden", "kat.json")))
    codes, sigma, _ = oracle.remap(b"mississippi")
    assert codes.tolist() == kat["mississippi"]["remapped"] and sigma == 5
    sa = oracle.sa(codes)
    assert oracle.c_table(codes, sigma).tolist() == kat["mississippi"]["c_table"]
    o = oracle.o_table(oracle.bwt(codes, sa), sigma)
    assert o.T.tolist() == kat["mississippi"]["o_rows_by_symbol"]
    codes, sigma, _ = oracle.remap(b"ababacabac")
    assert oracle.sa(codes).tolist() == kat["ababacabac"]["sa"]
    assert oracle.sa(codes, "qsort").tolist() == kat["ababacabac"]["sa"]
    codes, _, _ = oracle.remap(b"acagtgtaac")
    assert codes.tolist() == kat["acagtgtaac"]["remapped"]


def test_oracle_matches_golden(oracle, golden):
    for name in golden_cases(golden):
        codes = golden[f"{name}/codes"]
        sigma = int(golden[f"{name}/sigma"][0])
        sa = oracle.sa(codes)
        assert np.array_equal(sa, golden[f"{name}/sa"]), name
        if len(codes) <= 6000:
            assert np.array_equal(oracle.sa(codes, "qsort"), sa), name
        isa = oracle.inverse(sa)
        assert np.array_equal(isa, golden[f"{name}/isa"]), name
        if name != "empty_0":
            assert np.array_equal(oracle.lcp(codes, sa, isa), golden[f"{name}/lcp"]), name
        if f"{name}/c" in golden.files:
            assert np.array_equal(oracle.c_table(codes, sigma), golden[f"{name}/c"]), name
        bwt = oracle.bwt(codes, sa)
        if f"{name}/o" in golden.files:
            o = oracle.o_table(bwt, sigma)
            assert np.array_equal(o, golden[f"{name}/o"]), name
            ck = oracle.o_checkpoints(bwt, sigma, 64)
            assert np.array_equal(ck, o[::64]), name
        pats = sorted({k.split("/")[1] for k in golden.files if k.startswith(name + "/pat")})
        for p in pats:
            pc = golden[f"{name}/{p}/codes"]
            off = np.array([0, len(pc)], dtype=np.uint64)
            c = oracle.c_table(codes, sigma)
            ck = oracle.o_checkpoints(bwt, sigma, 64)
            L, R = oracle.search_ck(c, bwt, ck, 64, pc, off, threads=2)
            assert [int(L[0]), int(R[0])] == golden[f"{name}/{p}/LR"].tolist(), (name, p)
            _, pos = oracle.locate(sa, L, R)
            assert np.array_equal(pos, golden[f"{name}/{p}/pos"]), (name, p)
            if f"{name}/o" in golden.files:
                L2, R2 = oracle.search_dense(c, golden[f"{name}/o"], len(sa), pc, off)
                assert (L2[0], R2[0]) == (L[0], R[0])


@pytest.mark.parametrize("n,nsym", [(1, 1), (2, 4), (777, 4), (20000, 4), (20000, 2), (15000, 255), (30000, 1)])
def test_oracle_vs_reference_live(oracle, ref, n, nsym):
    if ref is None:
        pytest.skip("oracle/_ref/libstralg_ref.so not present")
    codes = oracle.random_codes(n, nsym, seed=n * 31 + nsym)
    sa_ref, isa_ref, lcp_ref = ref.sa_lcp(codes, nsym + 1)
    sa = oracle.sa(codes)
    assert np.array_equal(sa, sa_ref)
    assert np.array_equal(oracle.inverse(sa), isa_ref)
    assert np.array_equal(oracle.lcp(codes, sa), lcp_ref)


def test_oracle_search_vs_reference_iterator(oracle, ref):
    if ref is None:
        pytest.skip("oracle/_ref/libstralg_ref.so not present")
    rng = np.random.default_rng(5)
    raw = bytes(rng.choice(list(b"ACGT"), 50000).astype(np.uint8))
    t = ref.tables(raw)
    codes, sigma = t["codes"], t["sigma"]
    sa = oracle.sa(codes)
    assert np.array_equal(sa, t["sa"])
    bwt = oracle.bwt(codes, sa)
    c = oracle.c_table(codes, sigma)
    assert np.array_equal(c, t["c"])
    assert np.array_equal(oracle.o_table(bwt, sigma), t["o"])
    ck = oracle.o_checkpoints(bwt, sigma, 64)
    for k in range(200):
        m = int(rng.integers(1, 14))
        if k % 2:
            start = int(rng.integers(0, len(raw) - m))
            pc = codes[start:start + m]
        else:
            pc = rng.integers(1, 5, m).astype(np.uint8)
        L, R, pos = ref.exact_matches(t["handle"], pc)
        Lo, Ro = oracle.search_ck(c, bwt, ck, 64, pc, np.array([0, m], dtype=np.uint64))
        assert (int(Lo[0]), int(Ro[0])) == (L, R)
        _, po = oracle.locate(sa, Lo, Ro)
        assert np.array_equal(po, pos)
    ref.free_tables(t["handle"])


def test_synth_generators_are_deterministic(oracle):
    import ctypes as C
    a = np.empty(1001, dtype=np.uint8)
    b = np.empty(1001, dtype=np.uint8)
    for buf in (a, b):
        oracle.lib.oracle_synth_codes(buf.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_uint64(1000), C.c_uint32(4),
                                      C.c_uint64(42))
    assert np.array_equal(a, b) and a[-1] == 0 and a[:-1].min() >= 1 and a[:-1].max() <= 4


def test_oracle_approx_vs_reference_iterator(oracle, ref):
    """Pins the restatement of bwt.c:226-382 to the unmodified reference: interval list, matched
    lengths and CIGARs, in report order, with and without the reverse (RO / D table) tables; and the
    D table only prunes (same result either way)."""
    if ref is None:
        pytest.skip("oracle/_ref/libstralg_ref.so not present")
    rng = np.random.default_rng(1)
    intervals = 0
    for trial in range(12):
        n = int(rng.integers(5, 400))
        nsym = int(rng.integers(1, 5))
        raw = bytes(rng.choice(list(b"acgt"[:nsym]), n).tolist())
        for rev in (True, False):
            t = ref.tables(raw, rev)
            for q in range(6):
                m = int(rng.integers(1, 12))
                if q % 2 and n > m:
                    s = int(rng.integers(0, n - m))
                    pc = t["codes"][s:s + m].copy()
                    if m > 2:
                        pc[int(rng.integers(0, m))] = 1 + int(rng.integers(0, t["sigma"] - 1))
                else:
                    pc = rng.integers(1, t["sigma"], m).astype(np.uint8) if t["sigma"] > 1 else np.ones(m, np.uint8)
                for d in (0, 1, 2):
                    L, R, ml, cig, hits = ref.approx_matches(t["handle"], pc, d)
                    L2, R2, ml2, cig2 = oracle.approx(t["c"], t["o"], t["ro"], t["len"], pc, d)
                    assert np.array_equal(L, L2) and np.array_equal(R, R2) and np.array_equal(ml, ml2) and cig == cig2
                    if rev:
                        L3, R3, ml3, cig3 = oracle.approx(t["c"], t["o"], None, t["len"], pc, d)
                        assert np.array_equal(L, L3) and np.array_equal(R, R3) and cig == cig3
                    # the iterator yields SA[L..R) of every interval in turn (bwt.c:384-401)
                    exp = [(int(t["sa"][i]), cig[k], int(ml[k])) for k in range(len(L)) for i in range(int(L[k]), int(R[k]))]
                    assert hits == exp
                    intervals += len(L)
            ref.free_tables(t["handle"])
    assert intervals > 1000


def test_oracle_approx_known_answers(oracle):
    """mississippi, pattern "ssi": zero edits is the exact interval with CIGAR 3M; one edit adds the
    insertions the reference reports, in its order."""
    codes, sigma, table = oracle.remap(b"mississippi")
    sa = oracle.sa(codes)
    bwt = oracle.bwt(codes, sa)
    c, o = oracle.c_table(codes, sigma), oracle.o_table(bwt, sigma)
    p = oracle.remap_pattern(table, b"ssi")
    L, R, ml, cig = oracle.approx(c, o, None, len(codes), p, 0)
    assert cig == ["3M"] and ml.tolist() == [3] and sorted(sa[int(L[0]):int(R[0])].tolist()) == [2, 5]
    L, R, ml, cig = oracle.approx(c, o, None, len(codes), p, 1)
    # (values confirmed against the reference iterator by the test above)
    assert cig == ["3M", "1I2M", "1M1I1M", "2M1I"] and ml.tolist() == [3, 2, 2, 2]
    assert L.tolist() == [10, 8, 8, 10] and R.tolist() == [12, 10, 10, 12]


def test_oracle_approx_matches_golden(oracle):
    """tests/golden/approx_golden.json (the reference's iterator on its own test strings, d = 0..2):
    the restatement reproduces interval list, matched lengths, CIGARs and every yielded hit, in order,
    with the D table (reverse O table) and without it."""
    import json
    cases = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "approx_golden.json")))
    assert len(cases) > 200
    tables = {}
    for c in cases:
        if c["text"] not in tables:
            codes, sigma, _ = oracle.remap(c["text"].encode())
            sa = oracle.sa(codes)
            rcodes = np.concatenate([codes[:-1][::-1], np.zeros(1, np.uint8)])
            tables[c["text"]] = (codes, sa, oracle.c_table(codes, sigma), oracle.o_table(oracle.bwt(codes, sa), sigma),
                                 oracle.o_table(oracle.bwt(rcodes, oracle.sa(rcodes)), sigma))
        codes, sa, ct, o, ro = tables[c["text"]]
        for use_ro in (ro, None):
            L, R, ml, cig = oracle.approx(ct, o, use_ro, len(codes), np.array(c["codes"], np.uint8), c["edits"])
            assert L.tolist() == c["L"] and R.tolist() == c["R"] and ml.tolist() == c["match_length"] and cig == c["cigars"], c
        hits = [[int(sa[i]), cig[k], int(ml[k])] for k in range(len(L)) for i in range(int(L[k]), int(R[k]))]
        assert hits == c["hits"], c
