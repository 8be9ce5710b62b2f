"""GPU parity tests: the CUDA path (through the C ABI of include/b200sa.h) against the oracle.

Integer work: every comparison is bit-exact (np.array_equal).  Expected values come from the
committed golden fixtures, from the unmodified reference in oracle/_ref when its prebuilt library
travelled with the snapshot, and otherwise from the restatement oracle.
"""
import os

import numpy as np
import pytest

from conftest import golden_cases

pytestmark = pytest.mark.gpu


def expected_sa(oracle, ref, codes, sigma):
    if ref is not None and len(codes) > 2:
        return ref.sa(codes, sigma, "sa_is")
    return oracle.sa(codes)


def texts():
    """(name, symbols without sentinel, sigma)"""
    rng = np.random.default_rng(777)
    out = []
    for n in (1, 2, 3, 31, 63, 64, 65, 255, 2047, 2048, 2049, 6143, 6144, 6145, 16383, 16384, 16385, 100003):
        out.append((f"dna_{n}", rng.integers(1, 5, n), 5))
    out.append(("bin_50000", rng.integers(1, 3, 50000), 3))
    out.append(("unary_sigma5_30000", np.full(30000, 3), 5))
    out.append(("unary_sigma2_70001", np.ones(70001, dtype=np.int64), 2))
    out.append(("sym16_40000", rng.integers(1, 17, 40000), 17))
    out.append(("sym20_40000", rng.integers(1, 21, 40000), 21))
    out.append(("byte_60000", rng.integers(1, 256, 60000), 256))
    out.append(("acgt_period4_65537", np.tile([1, 2, 3, 4], 16385)[:65537], 5))
    out.append(("period1000_200000", np.tile(rng.integers(1, 5, 1000), 200), 5))
    a, b = [1], [1, 2]
    while len(b) < 100000:
        a, b = b, b + a
    out.append(("fibonacci_100000", np.array(b[:100000]), 3))
    out.append(("dna_tail_poly_a", np.concatenate([rng.integers(1, 5, 5000), np.ones(300, dtype=np.int64)]), 5))
    # small repetitive texts: buckets fit one SM but whole bins hold equal keys (robust in-SM sort)
    out.append(("unary_sigma2_3000", np.ones(3000, dtype=np.int64), 2))
    out.append(("unary_sigma5_2500", np.full(2500, 2), 5))
    out.append(("period3_3500", np.tile([2, 1, 3], 1200)[:3500], 5))
    out.append(("period50_3900", np.tile(rng.integers(1, 5, 50), 80)[:3900], 5))
    # a short suffix ("A$", padded with A) ties with a few long suffixes inside a run of A
    out.append(("dna_short_tie", np.concatenate([rng.integers(1, 5, 3000), np.ones(14, dtype=np.int64),
                                                 rng.integers(2, 5, 2000), [2, 1]]), 5))
    out.append(("dna_short_tie2", np.concatenate([rng.integers(1, 5, 200000), np.ones(16, dtype=np.int64),
                                                  rng.integers(2, 5, 2000), [1, 1, 1]]), 5))
    out.append(("dna_1M", rng.integers(1, 5, 1 << 20), 5))
    out.append(("repeat_rich_1M", np.tile(rng.integers(1, 5, 40000), 27)[: (1 << 20) - 3], 5))
    return out


TEXTS = texts()


# how round 0 (the initial K-symbol sort) runs: the default MSD bucket sort, the same with tiny
# digits (three partition levels even on small texts) or many small buckets per tile, and the
# LSD radix passes (also the path of texts whose buckets exceed one SM) at both digit widths
ROUND0_MODES = {
    "msd": {},
    "msd_3level": {"B200SA_MSD_DMAX": "4", "B200SA_MSD_BB": "12"},
    "msd_avg64": {"B200SA_MSD_AVG": "64"},
    "lsd8": {"B200SA_ROUND0": "lsd", "B200SA_RADIX_BITS": "8"},
    "lsd10": {"B200SA_ROUND0": "lsd", "B200SA_RADIX_BITS": "10"},
    # equal K-symbol keys are never decided by the next 64 bits of text in round 0 (all of them reach the rounds)
    "msd_noext": {"B200SA_NO_EXT_TIEBREAK": "1"},
    "msd_all_radix": {"B200SA_SMALL_PATH": "0", "B200SA_NO_EXT_TIEBREAK": "1"},
    "lsd8_noext": {"B200SA_ROUND0": "lsd", "B200SA_RADIX_BITS": "8", "B200SA_NO_EXT_TIEBREAK": "1"},
    # chain offsets in every doubling round, however small the active set (default: only large ones)
    "msd_chain": {"B200SA_CHAIN_MIN_FRAC": "100000000", "B200SA_CHAIN_USE_FRAC": "100000000", "B200SA_NO_EXT_TIEBREAK": "1"},
    "lsd8_chain": {"B200SA_ROUND0": "lsd", "B200SA_RADIX_BITS": "8", "B200SA_CHAIN_MIN_FRAC": "100000000", "B200SA_CHAIN_USE_FRAC": "100000000"},
    # pivot path (sa_build.cu pivot_classify_kernel / pivot_apply_kernel) in every doubling round whatever it finds,
    # and with its own heuristics on lists of any size
    "msd_pivot": {"B200SA_PIVOT_MIN": "2", "B200SA_PIVOT_FORCE": "1", "B200SA_NO_EXT_TIEBREAK": "1"},
    "lsd8_pivot": {"B200SA_ROUND0": "lsd", "B200SA_RADIX_BITS": "8", "B200SA_PIVOT_MIN": "2", "B200SA_PIVOT_FORCE": "1"},
    "lsd10_pivot_auto": {"B200SA_ROUND0": "lsd", "B200SA_RADIX_BITS": "10", "B200SA_PIVOT_MIN": "2"},
    "msd_pivot_auto_nochain": {"B200SA_PIVOT_MIN": "2", "B200SA_CHAIN": "0", "B200SA_NO_EXT_TIEBREAK": "1"},
}


@pytest.mark.parametrize("mode", list(ROUND0_MODES))
@pytest.mark.parametrize("case", TEXTS, ids=[t[0] for t in TEXTS])
def test_suffix_array_tables_match_oracle(engine, oracle, ref, case, mode, monkeypatch):
    for k, v in ROUND0_MODES[mode].items():
        monkeypatch.setenv(k, v)
    name, sym, sigma = case
    codes = np.concatenate([np.asarray(sym, dtype=np.uint8), np.zeros(1, np.uint8)])
    idx = engine.SuffixArrayIndex.build(codes[:-1], sigma, isa=True, lcp=True, bwt=True, occ=True)
    sa_exp = expected_sa(oracle, ref, codes, sigma)
    sa = idx.sa()
    if name.startswith(("dna_", "sym", "byte_", "bin_")):
        # random texts must really take the path under test (no silent fallback)
        assert idx.stats()["round0_mode"] == (1 if mode.startswith("msd") else 0), (name, idx.stats())
    assert np.array_equal(sa, sa_exp), f"{name}: first mismatch at {np.nonzero(sa != sa_exp)[0][:5]}"
    assert np.array_equal(idx.isa(), oracle.inverse(sa_exp))
    lcp_exp = oracle.lcp(codes, sa_exp)
    lcp = idx.lcp()
    assert np.array_equal(lcp, lcp_exp), f"{name}: LCP mismatch at {np.nonzero(lcp != lcp_exp)[0][:5]}"
    bwt_exp = oracle.bwt(codes, sa_exp)
    assert np.array_equal(idx.bwt(), bwt_exp)
    assert np.array_equal(idx.c_table(), oracle.c_table(codes, sigma))
    assert idx.primary == int(np.nonzero(sa_exp == 0)[0][0])
    # O(a, i): every checkpoint row plus random probes; the full dense table for small inputs
    ck = oracle.o_checkpoints(bwt_exp, sigma, 64)
    rows = np.arange(0, len(sa) + 1, 64, dtype=np.uint32)
    rng = np.random.default_rng(1)
    for a in sorted(set([0, 1, sigma - 1, int(rng.integers(0, sigma))])):
        got = idx.occ(np.full(len(rows), a, dtype=np.uint8), rows)
        assert np.array_equal(got, ck[:, a]), (name, a)
    qa = rng.integers(0, sigma, 5000).astype(np.uint8)
    qi = rng.integers(0, len(sa) + 1, 5000).astype(np.uint32)
    qi[:4] = [0, len(sa), len(sa) - 1, min(1, len(sa))]
    assert np.array_equal(idx.occ(qa, qi), oracle.o_probe(bwt_exp, ck, sigma, 64, qa, qi))
    if len(sa) <= 20000 and sigma <= 21:
        assert np.array_equal(idx.o_dense(), oracle.o_table(bwt_exp, sigma))
    idx.close()


def test_golden_fixtures(engine, golden):
    """Outputs of the reference itself on its own test strings (tests/golden/make_golden.py)."""
    for name in golden_cases(golden):
        codes = golden[f"{name}/codes"]
        sigma = int(golden[f"{name}/sigma"][0])
        idx = engine.SuffixArrayIndex.build(codes[:-1], sigma, isa=True, lcp=True, bwt=True, occ=True)
        assert np.array_equal(idx.sa(), golden[f"{name}/sa"]), name
        assert np.array_equal(idx.isa(), golden[f"{name}/isa"]), name
        assert np.array_equal(idx.lcp(), golden[f"{name}/lcp"]), name
        if f"{name}/c" in golden.files:
            assert np.array_equal(idx.c_table(), golden[f"{name}/c"]), name
        if f"{name}/o" in golden.files:
            assert np.array_equal(idx.o_dense(), golden[f"{name}/o"]), name
        pats = sorted({k.split("/")[1] for k in golden.files if k.startswith(name + "/pat")})
        for p in pats:
            pc = golden[f"{name}/{p}/codes"]
            L, R = idx.search_one(pc)
            assert [L, R] == golden[f"{name}/{p}/LR"].tolist(), (name, p)
            assert np.array_equal(idx.exact_matches(pc), golden[f"{name}/{p}/pos"]), (name, p)
        idx.close()


def make_patterns(rng, codes, nsym, npat, mmin, mmax):
    n = len(codes) - 1
    lens = rng.integers(mmin, mmax + 1, npat)
    off = np.zeros(npat + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    pat = np.empty(int(off[-1]), dtype=np.uint8)
    for k in range(npat):
        m = int(lens[k])
        if k % 2 and n > m:
            s = int(rng.integers(0, n - m))
            pat[int(off[k]):int(off[k + 1])] = codes[s:s + m]
        else:
            pat[int(off[k]):int(off[k + 1])] = rng.integers(1, nsym + 1, m)
    return pat, off


@pytest.mark.parametrize("textcmp,ktable", [(False, False), (True, False), (False, True), (True, True)],
                         ids=["plain", "textcmp", "ktable", "textcmp_ktable"])
@pytest.mark.parametrize("n,nsym,mmax", [(1 << 20, 4, 24), (200000, 4, 40), (50000, 2, 30), (80000, 20, 6),
                                         (60000, 255, 4), (3000, 1, 50), (40000, 3, 200), (5000, 4, 300),
                                         # other alphabets with patterns long enough for the single-candidate text
                                         # comparison of the generic kernel (fm_search.cu: fm_search_kernel<2, true>)
                                         (100000, 5, 60), (80000, 20, 40), (30000, 100, 30), (20000, 6, 120)])
def test_batched_search_and_locate(engine, oracle, n, nsym, mmax, textcmp, ktable):
    """(L, R) per pattern and the position sets against the restatement of bwt.c:164-217;
    config 1 of BASELINE.json is the first row (1 Mi random ACGT, 10 k patterns)."""
    rng = np.random.default_rng(n + nsym)
    codes = oracle.random_codes(n, nsym, seed=n)
    sigma = nsym + 1
    idx = engine.SuffixArrayIndex.build(codes[:-1], sigma, textcmp=textcmp, ktable=ktable)
    sa = idx.sa()
    bwt = oracle.bwt(codes, sa)
    ck = oracle.o_checkpoints(bwt, sigma, 64)
    c = oracle.c_table(codes, sigma)
    pat, off = make_patterns(rng, codes, nsym, 10000, 1, mmax)
    # near-misses: copies of text windows with one symbol changed (exercise the failing step)
    for k in range(0, 10000, 7):
        lo, hi = int(off[k]), int(off[k + 1])
        if hi - lo >= 2 and k % 2:
            j = lo + int(rng.integers(0, hi - lo))
            pat[j] = 1 + (pat[j] % nsym)
    if ktable and n >= 20000 and nsym <= 20:
        assert idx.stats()["ktable_k"] >= 2, idx.stats()  # (every alphabet whose k-mers fit the budget has a seed table)
    L, R = idx.search(pat, off)
    Le, Re = oracle.search_ck(c, bwt, ck, 64, pat, off, threads=4)
    assert np.array_equal(L, Le) and np.array_equal(R, Re)
    poff, pos = idx.locate(L, R)
    poff_e, pos_e = oracle.locate(oracle.sa(codes) if n <= 3000 else sa, Le, Re)
    assert np.array_equal(poff, poff_e) and np.array_equal(pos, pos_e)
    # fixed-length batch entry point
    m = min(20, n)
    fixed = np.concatenate([codes[s:s + m] for s in rng.integers(0, n - m + 1, 500)])
    Lf, Rf = idx.search(fixed, fixed_len=m)
    Lfe, Rfe = oracle.search_ck(c, bwt, ck, 64, fixed, np.arange(0, 501 * m, m, dtype=np.uint64))
    assert np.array_equal(Lf, Lfe) and np.array_equal(Rf, Rfe) and (Rf > Lf).all()
    idx.close()


def test_search_edge_cases(engine, oracle):
    codes, sigma, table = oracle.remap(b"mississippi")
    for textcmp in (False, True):
        _search_edge_cases(engine, codes, sigma, textcmp)


def _search_edge_cases(engine, codes, sigma, textcmp):
    idx = engine.SuffixArrayIndex.build(codes[:-1], sigma, textcmp=textcmp)
    # pattern longer than the text: (1, 0) like bwt.c:179-181
    long_pat = np.tile(codes[:-1], 2)
    assert idx.search_one(long_pat) == (1, 0)
    # a symbol outside 1..sigma-1: empty interval
    assert idx.search_one(np.array([1, 7], dtype=np.uint8)) == (1, 0)
    # whole text matches once at position 0
    L, R = idx.search_one(codes[:-1])
    assert R - L == 1 and idx.exact_matches(codes[:-1]).tolist() == [0]
    idx.close()


def test_bad_symbol_is_reported(engine):
    with pytest.raises(engine.B200saError) as e:
        engine.SuffixArrayIndex.build(np.array([1, 2, 0, 3], dtype=np.uint8), 5)
    assert e.value.code == 3
    with pytest.raises(engine.B200saError) as e:
        engine.SuffixArrayIndex.build(np.array([1, 2, 5, 3], dtype=np.uint8), 5)
    assert e.value.code == 3


def test_device_resident_text_and_synth(engine, oracle):
    """Device-pointer entry points: text generated on the GPU, built without a host copy."""
    import ctypes as C
    import torch
    n = 300000
    text = torch.empty(n + 1, dtype=torch.uint8, device="cuda")
    lib = engine.load()
    assert lib.b200sa_synth_codes(C.c_void_p(text.data_ptr()), n, 4, 99, 0, None) == 0
    torch.cuda.synchronize()
    host = np.empty(n + 1, dtype=np.uint8)
    oracle.lib.oracle_synth_codes(host.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_uint64(n), C.c_uint32(4),
                                  C.c_uint64(99))
    assert np.array_equal(text.cpu().numpy(), host)
    idx = engine.SuffixArrayIndex.build(text[:n], 5)
    assert np.array_equal(idx.sa(), oracle.sa(host))
    reads = torch.empty(1000 * 30, dtype=torch.uint8, device="cuda")
    assert lib.b200sa_synth_reads(C.c_void_p(text.data_ptr()), n, 4, C.c_void_p(reads.data_ptr()), 1000, 30, 100, 7,
                                  0, None) == 0
    hreads = np.empty(1000 * 30, dtype=np.uint8)
    oracle.lib.oracle_synth_reads(host.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_uint64(n), C.c_uint32(4),
                                  hreads.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_uint64(1000), C.c_uint32(30),
                                  C.c_uint32(100), C.c_uint64(7))
    torch.cuda.synchronize()
    assert np.array_equal(reads.cpu().numpy(), hreads)
    dL = torch.empty(1000, dtype=torch.int32, device="cuda")
    dR = torch.empty(1000, dtype=torch.int32, device="cuda")
    idx.search_device(reads, None, 30, 1000, dL, dR, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    L, R = idx.search(hreads, fixed_len=30)
    assert np.array_equal(dL.cpu().numpy().view(np.uint32), L) and np.array_equal(dR.cpu().numpy().view(np.uint32), R)
    assert ((R > L).sum()) >= 850  # ~90 % of the reads are sampled from the text
    idx.close()
    # the counting variant of the search kernel (b200sa_search_traffic) returns the same intervals and
    # plausible operation counts, with and without the unique-interval shortcut
    for textcmp in (False, True):
        idx = engine.SuffixArrayIndex.build(text[:n], 5, textcmp=textcmp)
        counts = (C.c_uint64 * 4)()
        dL.zero_()
        dR.zero_()
        rc = lib.b200sa_search_traffic(idx._h, C.c_void_p(reads.data_ptr()), None, 30, 1000, C.c_void_p(dL.data_ptr()),
                                       C.c_void_p(dR.data_ptr()), counts, None)
        assert rc == 0
        assert np.array_equal(dL.cpu().numpy().view(np.uint32), L) and np.array_equal(dR.cpu().numpy().view(np.uint32), R)
        blocks, pwords, twords, sa_isa = [int(x) for x in counts]
        assert 1000 * 5 <= blocks <= 1000 * 60 and 1000 <= pwords <= 1000 * 7
        if textcmp:
            assert twords > 0 and sa_isa > 0 and blocks < 1000 * 40
        else:
            assert twords == 0 and sa_isa == 0
        idx.close()


def test_host_batch_pipelined_pieces_equal_single_launch(engine, oracle):
    """b200sa_search_batch cuts a large host batch into pieces that alternate between two streams
    (copies overlap the kernels); the result must equal one launch over device-resident reads.
    140 MB of patterns: fixed-length and variable-length (absolute offsets) entry points."""
    import ctypes as C
    import torch
    n, m, npat = 1 << 20, 100, 1_400_000
    codes = oracle.random_codes(n, 4, seed=99)
    idx = engine.SuffixArrayIndex.build(codes[:-1], 5, textcmp=True, ktable=True)
    lib = engine.load()
    text = torch.from_numpy(codes).cuda()
    reads = torch.empty(npat * m, dtype=torch.uint8, device="cuda")
    assert lib.b200sa_synth_reads(C.c_void_p(text.data_ptr()), n, 4, C.c_void_p(reads.data_ptr()), npat, m, 102, 7,
                                  0, None) == 0
    Ld = torch.empty(npat, dtype=torch.int32, device="cuda")
    Rd = torch.empty(npat, dtype=torch.int32, device="cuda")
    idx.search_device(reads, None, m, npat, Ld, Rd)
    torch.cuda.synchronize()
    Le, Re = Ld.cpu().numpy().view(np.uint32), Rd.cpu().numpy().view(np.uint32)
    h = reads.cpu().numpy()
    L, R = idx.search(h, fixed_len=m)
    assert np.array_equal(L, Le) and np.array_equal(R, Re) and (R > L).mean() > 0.8
    # variable lengths: drop a different number of leading symbols from every read
    cut = (np.arange(npat) % 7).astype(np.uint64)
    off = np.zeros(npat + 1, dtype=np.uint64)
    off[1:] = np.cumsum(m - cut)
    keep = np.ones(npat * m, dtype=bool)
    for c in range(1, 7):
        rows = np.nonzero(cut == c)[0]
        for j in range(c):
            keep[rows * m + j] = False
    hv = h[keep]
    assert hv.size == int(off[-1])
    Lv, Rv = idx.search(hv, off)
    dv, doff = torch.from_numpy(hv).cuda(), torch.from_numpy(off.view(np.int64)).cuda()
    idx.search_device(dv, doff, 0, npat, Ld, Rd)
    torch.cuda.synchronize()
    assert np.array_equal(Lv, Ld.cpu().numpy().view(np.uint32)) and np.array_equal(Rv, Rd.cpu().numpy().view(np.uint32))
    idx.close()


@pytest.mark.parametrize("textcmp,ktable", [(False, False), (True, False), (False, True), (True, True)],
                         ids=["plain", "textcmp", "ktable", "textcmp_ktable"])
@pytest.mark.parametrize("n,m,stride", [(1 << 20, 100, 0), (1 << 20, 101, 32), (300000, 37, 0), (300000, 3, 0),
                                        (50000, 1, 0), (200000, 250, 64), (3000, 64, 0), (1 << 20, 32, 8)])
def test_packed_reads_equal_byte_reads(engine, oracle, n, m, stride, textcmp, ktable):
    """b200sa_search_batch_packed / _device_packed (2 bits per base) return the (L, R)
    of b200sa_search_batch on the same reads (bwt.c:164-199), hits, misses and near-misses alike; the
    restatement pins the byte API in test_batched_search_and_locate."""
    import ctypes as C
    import torch
    rng = np.random.default_rng(n + m)
    codes = oracle.random_codes(n, 4, seed=n + 1)
    idx = engine.SuffixArrayIndex.build(codes[:-1], 5, textcmp=textcmp, ktable=ktable)
    npat = 20000
    starts = rng.integers(0, n - m + 1, npat)
    reads = np.stack([codes[s:s + m] for s in starts]).astype(np.uint8)
    reads[::5] = rng.integers(1, 5, (len(reads[::5]), m))            # random reads: misses
    for q in range(1, npat, 3):                                      # near-misses: one base changed
        j = int(rng.integers(0, m))
        reads[q, j] = 1 + reads[q, j] % 4
    reads[7] = codes[:m]                                             # the text's first bases (suffix 0)
    reads[8] = codes[n - m:n]                                        # its last bases
    flat = reads.reshape(-1)
    L, R = idx.search(flat, fixed_len=m)
    if m >= 20:
        assert 0.2 < float((R > L).mean()) < 0.9
    packed = engine.pack_reads(flat, m, stride)
    Lp, Rp = idx.search_packed(packed, m, npat, stride)
    assert np.array_equal(Lp, L) and np.array_equal(Rp, R), np.nonzero((Lp != L) | (Rp != R))[0][:10]
    # device entry points: pack on the device, search there
    lib = engine.load()
    d_codes = torch.from_numpy(flat).cuda()
    st = stride or (m + 3) // 4
    d_packed = torch.zeros(npat * st + 8, dtype=torch.uint8, device="cuda")
    assert lib.b200sa_pack_reads_device(C.c_void_p(d_codes.data_ptr()), m, st, npat, C.c_void_p(d_packed.data_ptr()),
                                        0, None) == 0
    assert np.array_equal(d_packed.cpu().numpy(), packed[: npat * st + 8])
    dL = torch.empty(npat, dtype=torch.int32, device="cuda")
    dR = torch.empty(npat, dtype=torch.int32, device="cuda")
    idx.search_device_packed(d_packed, m, npat, dL, dR, st)
    torch.cuda.synchronize()
    assert np.array_equal(dL.cpu().numpy().view(np.uint32), L) and np.array_equal(dR.cpu().numpy().view(np.uint32), R)
    counts = (C.c_uint64 * 4)()
    assert lib.b200sa_search_traffic_packed(idx._h, C.c_void_p(d_packed.data_ptr()), m, st, npat,
                                            C.c_void_p(dL.data_ptr()), C.c_void_p(dR.data_ptr()), counts, None) == 0
    assert np.array_equal(dL.cpu().numpy().view(np.uint32), L) and counts[0] > 0 and counts[1] > 0
    idx.close()


def test_packed_reads_reject_bad_input(engine, oracle):
    codes = oracle.random_codes(5000, 4, seed=3)
    idx = engine.SuffixArrayIndex.build(codes[:-1], 5)
    with pytest.raises(engine.B200saError) as e:
        engine.pack_reads(np.array([1, 2, 5, 1], dtype=np.uint8), 4)
    assert e.value.code == 3
    with pytest.raises(engine.B200saError) as e:
        idx.search_packed(np.zeros(16, np.uint8), 8, 1, stride_bytes=1)
    assert e.value.code == 2
    idx.close()
    big = engine.SuffixArrayIndex.build(oracle.random_codes(5000, 20, seed=3)[:-1], 21)
    with pytest.raises(engine.B200saError) as e:
        big.search_packed(np.zeros(16, np.uint8), 8, 1)
    assert e.value.code == 5
    big.close()


def test_one_pattern_mailbox_path(engine, oracle):
    """A handful of short patterns take the mapped-memory path of b200sa_search_batch (one launch, no
    cudaMemcpy): same intervals as a large batch of the same patterns."""
    rng = np.random.default_rng(12)
    n = 200000
    codes = oracle.random_codes(n, 4, seed=77)
    for textcmp, ktable in ((False, False), (True, True)):
        idx = engine.SuffixArrayIndex.build(codes[:-1], 5, textcmp=textcmp, ktable=ktable)
        pats, lens = [], rng.integers(1, 120, 300)
        for k, m in enumerate(lens):
            s = int(rng.integers(0, n - m))
            p = codes[s:s + m].copy()
            if k % 3 == 0:
                p[int(rng.integers(0, m))] = 1 + int(rng.integers(0, 4))
            pats.append(p)
        off = np.zeros(len(pats) + 1, dtype=np.uint64)
        off[1:] = np.cumsum(lens)
        Lb, Rb = idx.search(np.concatenate(pats), off)
        for k, p in enumerate(pats):
            assert idx.search_one(p) == (int(Lb[k]), int(Rb[k])), k
        # up to 64 patterns at once, offsets not starting at 0
        L8, R8 = idx.search(np.concatenate(pats[:8]), off[:9])
        assert np.array_equal(L8, Lb[:8]) and np.array_equal(R8, Rb[:8])
        L9, R9 = idx.search(np.concatenate(pats), off[5:14])
        assert np.array_equal(L9, Lb[5:13]) and np.array_equal(R9, Rb[5:13])
        idx.close()


def test_one_pattern_at_a_time_resident_kernel(engine, oracle):
    """The drop-in's one-iterator-per-pattern use (bwt.c:164-199): single-pattern calls are served by a resident
    one-warp kernel that polls a mailbox in pinned memory (fm_search.cu fm_mailbox_server_kernel).  Same (L, R) as the
    batch kernel -- hits, misses, near-misses, a bad symbol, the empty pattern, lengths up to the mailbox limit and
    beyond it (launch path) -- across the life cycle of the kernel: a second index taking the mailbox over, an index
    being extended and freed while the kernel serves it, a pause longer than its idle limit."""
    import time
    rng = np.random.default_rng(123)
    n = 300000
    codes_a = oracle.random_codes(n, 4, seed=11)
    codes_b = oracle.random_codes(n // 2, 4, seed=12)
    ia = engine.SuffixArrayIndex.build(codes_a[:-1], 5, textcmp=True, ktable=True)
    ib = engine.SuffixArrayIndex.build(codes_b[:-1], 5)  # plain recurrence: no seed table, no text comparison

    def patterns(codes, count):
        out = []
        for _ in range(count):
            m = int(rng.choice([1, 2, 7, 20, 33, 100, 150, 186, 187, 250]))
            s = int(rng.integers(0, len(codes) - 1 - m))
            p = codes[s:s + m].copy()
            kind = rng.integers(0, 4)
            if kind == 1:
                p[rng.integers(0, m)] = rng.integers(1, 5)      # near-miss
            elif kind == 2:
                p = rng.integers(1, 5, m).astype(np.uint8)        # random
            elif kind == 3 and m > 2:
                p[rng.integers(0, m)] = 7                          # a symbol the text does not have
            out.append(p.astype(np.uint8))
        out.append(np.zeros(0, np.uint8))
        return out

    def check(idx, codes, count):
        pats = patterns(codes, count)
        off = np.concatenate([[0], np.cumsum([len(p) for p in pats])]).astype(np.uint64)
        cat = np.concatenate(pats + [np.zeros(16, np.uint8)])
        Lb, Rb = idx.search(cat, off)  # batch kernel
        for q, p in enumerate(pats):
            L1, R1 = idx.search(np.concatenate([p, np.zeros(8, np.uint8)]), np.array([0, len(p)], dtype=np.uint64))
            assert (int(L1[0]), int(R1[0])) == (int(Lb[q]), int(Rb[q])), (q, len(p), L1, R1, Lb[q], Rb[q])

    check(ia, codes_a, 300)
    check(ib, codes_b, 200)          # the mailbox changes hands
    check(ia, codes_a, 50)
    time.sleep(0.05)                  # longer than the idle limit: the kernel has left and is launched again
    check(ia, codes_a, 50)
    ib2 = engine.SuffixArrayIndex.build(codes_b[:-1], 5)
    check(ib2, codes_b, 20)
    ib2.extend(codes_b[:-1], textcmp=True, ktable=True)  # tables change under the resident kernel: it is stopped first
    check(ib2, codes_b, 100)
    ib2.close()                       # freed while it holds the mailbox
    check(ia, codes_a, 50)
    ia.close()
    ib.close()
