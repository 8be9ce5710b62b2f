"""GPU parity tests for the batched approximate search against the oracle's
restatement of stralg/bwt.c:226-382 and, when its prebuilt library travelled with the snapshot,
against the unmodified reference iterator.  Everything is compared exactly and IN ORDER: interval
list (L, R), matched lengths, CIGAR strings, and the positions the iterator yields.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _dense_tables(oracle, codes, sigma):
    sa = oracle.sa(codes)
    bwt = oracle.bwt(codes, sa)
    return sa, oracle.c_table(codes, sigma), oracle.o_table(bwt, sigma)


def _reverse_codes(codes):
    return np.concatenate([codes[:-1][::-1], np.zeros(1, np.uint8)])


def _patterns(rng, codes, nsym, npat, mmin, mmax):
    n = len(codes) - 1
    pats = []
    for k in range(npat):
        m = int(rng.integers(mmin, mmax + 1))
        if k % 3 and n > m:
            s = int(rng.integers(0, n - m))
            p = codes[s:s + m].copy()
            for _ in range(k % 3 - 1 + int(rng.integers(0, 2))):   # a few edits
                j = int(rng.integers(0, m))
                p[j] = 1 + int(rng.integers(0, nsym))
        else:
            p = rng.integers(1, nsym + 1, m).astype(np.uint8)
        pats.append(p)
    off = np.zeros(npat + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(p) for p in pats])
    return np.concatenate(pats).astype(np.uint8), off, pats


CASES = [  # n, nsym, npat, mmin, mmax
    (3000, 4, 300, 1, 14),
    (20000, 4, 300, 8, 25),
    (500, 2, 200, 1, 10),
    (64, 4, 100, 1, 8),
    (2000, 1, 50, 1, 8),        # unary text
    (4000, 20, 200, 1, 7),      # byte-block O layout
    (1500, 200, 100, 1, 5),
]


@pytest.mark.parametrize("use_rev", [True, False], ids=["dtable", "no_dtable"])
@pytest.mark.parametrize("n,nsym,npat,mmin,mmax", CASES)
def test_approx_matches_oracle(engine, oracle, n, nsym, npat, mmin, mmax, use_rev):
    rng = np.random.default_rng(n + 7 * nsym)
    codes = oracle.random_codes(n, nsym, seed=n + nsym)
    sigma = nsym + 1
    sa, c, o = _dense_tables(oracle, codes, sigma)
    rcodes = _reverse_codes(codes)
    _, _, ro = _dense_tables(oracle, rcodes, sigma)
    idx = engine.SuffixArrayIndex.build(codes[:-1], sigma)
    rev = engine.SuffixArrayIndex.build(rcodes[:-1], sigma, drop_sa=True) if use_rev else None
    pat, off, pats = _patterns(rng, codes, nsym, npat, mmin, mmax)
    for d in (0, 1, 2):
        if d == 2 and nsym > 4:
            continue  # (the oracle's walk is the slow side here)
        res = idx.approx_search(pat, off, max_edits=d, rev=rev)
        eL, eR, eM, eC, eoff = [], [], [], [], [0]
        for p in pats:
            L, R, ml, cig = oracle.approx(c, o, ro if use_rev else None, len(codes), p, d)
            eL.append(L); eR.append(R); eM.append(ml); eC += cig
            eoff.append(eoff[-1] + len(L))
        assert np.array_equal(res["offsets"], np.array(eoff, dtype=np.uint64)), d
        assert np.array_equal(res["L"], np.concatenate(eL)) and np.array_equal(res["R"], np.concatenate(eR)), d
        assert np.array_equal(res["match_length"], np.concatenate(eM)), d
        assert res["cigars"] == eC, d
        # positions: the iterator walks every interval in turn (bwt.c:384-401)
        _, pos = idx.locate(res["L"], res["R"])
        _, pos_e = oracle.locate(sa, np.concatenate(eL).astype(np.uint32), np.concatenate(eR).astype(np.uint32))
        assert np.array_equal(pos, pos_e)
    if rev is not None:
        # a precomputed D table gives the same answer as the reverse index
        dt = np.concatenate([oracle.approx_dtable(c, ro, len(codes), p) for p in pats]).astype(np.uint8)
        a = idx.approx_search(pat, off, max_edits=1, d_table=dt)
        b = idx.approx_search(pat, off, max_edits=1, rev=rev)
        assert all(np.array_equal(a[k], b[k]) for k in ("offsets", "L", "R", "match_length")) and a["cigars"] == b["cigars"]
        rev.close()
    idx.close()


def test_approx_matches_reference_iterator(engine, oracle, ref):
    """The reference's own iterator on its own table (build_complete_table with the reverse
    tables): interval list, CIGARs and every (position, cigar, match_length) it yields."""
    if ref is None:
        pytest.skip("oracle/_ref/libstralg_ref.so did not travel with this snapshot")
    rng = np.random.default_rng(99)
    texts = [b"mississippi", b"ababacabac", b"acagtgtaac", b"aaaaaaaaaaaaaaaa",
             bytes(rng.choice(list(b"acgt"), 700).tolist())]
    for raw in texts:
        t = ref.tables(raw, include_reverse=True)
        codes, sigma = t["codes"], t["sigma"]
        idx = engine.SuffixArrayIndex.build(codes[:-1], sigma)
        rev = engine.SuffixArrayIndex.build(codes[:-1][::-1].copy(), sigma, drop_sa=True)
        pats = [codes[s:s + m].copy() for s in range(0, len(raw) - 1, max(1, len(raw) // 9)) for m in (1, 3, 6)
                if s + m <= len(raw)]
        pats += [rng.integers(1, sigma, int(rng.integers(1, 8))).astype(np.uint8) if sigma > 1 else np.ones(3, np.uint8)
                 for _ in range(20)]
        off = np.zeros(len(pats) + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(p) for p in pats])
        pat = np.concatenate(pats).astype(np.uint8)
        for d in (0, 1, 2):
            res = idx.approx_search(pat, off, max_edits=d, rev=rev)
            poff, pos = idx.locate(res["L"], res["R"])
            for q, p in enumerate(pats):
                L, R, ml, cig, hits = ref.approx_matches(t["handle"], p, d)
                a, b = int(res["offsets"][q]), int(res["offsets"][q + 1])
                assert np.array_equal(res["L"][a:b], L) and np.array_equal(res["R"][a:b], R), (raw, p, d)
                assert np.array_equal(res["match_length"][a:b], ml) and res["cigars"][a:b] == cig, (raw, p, d)
                mine = [(int(pos[k]), res["cigars"][h], int(res["match_length"][h]))
                        for h in range(a, b) for k in range(int(poff[h]), int(poff[h + 1]))]
                assert mine == hits, (raw, p, d)
        ref.free_tables(t["handle"])
        idx.close()
        rev.close()


def test_approx_edge_cases(engine, oracle):
    codes, sigma, table = oracle.remap(b"mississippi")
    idx = engine.SuffixArrayIndex.build(codes[:-1], sigma)
    # no patterns
    r = idx.approx_search(np.zeros(0, np.uint8), np.zeros(1, np.uint64), max_edits=1)
    assert r["offsets"].tolist() == [0] and r["cigars"] == []
    # zero edits == exact search: one interval, CIGAR "mM"
    p = oracle.remap_pattern(table, b"ssi")
    r = idx.approx_search(p, np.array([0, 3], np.uint64), max_edits=0)
    assert (int(r["L"][0]), int(r["R"][0])) == idx.search_one(p) and r["cigars"] == ["3M"]
    # an empty pattern among others yields nothing (the reference asserts m > 0, bwt.c:343)
    r = idx.approx_search(p, np.array([0, 0, 3], np.uint64), max_edits=0)
    assert r["offsets"].tolist() == [0, 0, 1]
    with pytest.raises(engine.B200saError):
        idx.approx_search(p, np.array([0, 3], np.uint64), max_edits=300)
    idx.close()


def test_approx_matches_golden(engine, oracle):
    """tests/golden/approx_golden.json: the reference iterator's output on its own test strings
    (match_test.c:682-696, mississippi) at d = 0, 1, 2 -- interval list, matched lengths, CIGARs and every
    yielded (position, CIGAR, length), in order; with the D table from a reverse index and without."""
    import json
    import os
    cases = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "approx_golden.json")))
    by_text = {}
    for c in cases:
        by_text.setdefault(c["text"], []).append(c)
    for text, cs in by_text.items():
        codes, sigma, _ = oracle.remap(text.encode())
        idx = engine.SuffixArrayIndex.build(codes[:-1], sigma)
        rev = engine.SuffixArrayIndex.build(codes[:-1][::-1].copy(), sigma, drop_sa=True)
        sa = idx.sa()
        for d in (0, 1, 2):
            group = [c for c in cs if c["edits"] == d]
            pat = np.concatenate([np.array(c["codes"], np.uint8) for c in group])
            off = np.zeros(len(group) + 1, dtype=np.uint64)
            off[1:] = np.cumsum([len(c["codes"]) for c in group])
            for use_rev in (rev, None):
                res = idx.approx_search(pat, off, max_edits=d, rev=use_rev)
                for q, c in enumerate(group):
                    a, b = int(res["offsets"][q]), int(res["offsets"][q + 1])
                    assert res["L"][a:b].tolist() == c["L"] and res["R"][a:b].tolist() == c["R"], c
                    assert res["match_length"][a:b].tolist() == c["match_length"] and res["cigars"][a:b] == c["cigars"], c
                    hits = [[int(sa[i]), res["cigars"][h], int(res["match_length"][h])]
                            for h in range(a, b) for i in range(int(res["L"][h]), int(res["R"][h]))]
                    assert hits == c["hits"], c
        idx.close()
        rev.close()
