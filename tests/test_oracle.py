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
