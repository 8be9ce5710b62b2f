"""GPU parity tests for locate through the SAMPLED suffix array and for sorted positions
(reference semantics: next_bwt_exact_match_iter, stralg/bwt.c:201-217, and the
sorted comparison of tests/stralg/match_test.c:608).  Integer work: every comparison is bit-exact.
"""
import numpy as np
import pytest

from test_gpu_parity import make_patterns

pytestmark = pytest.mark.gpu


CASES = [  # n, nsym, rate, mmax
    (1 << 20, 4, 32, 24),
    (200000, 4, 7, 30),       # a rate that is not a power of two
    (100003, 4, 1, 12),       # every row sampled: zero LF steps
    (70001, 4, 64, 3),        # short patterns: wide intervals
    (50000, 2, 16, 30),
    (80000, 20, 8, 6),        # byte-block O layout
    (60000, 255, 5, 4),
    (3000, 1, 33, 50),        # unary text: every LF walk runs down the same chain
    (65, 4, 4, 5), (64, 4, 3, 5), (63, 4, 100, 5), (1, 4, 2, 1),
]


@pytest.mark.parametrize("n,nsym,rate,mmax", CASES)
def test_sampled_sa_locate_matches_full_sa(engine, oracle, n, nsym, rate, mmax):
    rng = np.random.default_rng(n * 31 + rate)
    codes = oracle.random_codes(n, nsym, seed=n + rate)
    sigma = nsym + 1
    idx = engine.SuffixArrayIndex.build(codes[:-1], sigma)
    sa = idx.sa()
    assert n > 200000 or np.array_equal(sa, oracle.sa(codes))  # (large n: checked in test_gpu_parity)
    pat, off = make_patterns(rng, codes, nsym, 4000, 1, mmax)
    L, R = idx.search(pat, off)
    poff_full, pos_full = idx.locate(L, R)
    poff_e, pos_e = oracle.locate(sa, L, R)
    assert np.array_equal(poff_full, poff_e) and np.array_equal(pos_full, pos_e)

    idx.sample_sa(rate, drop_sa=False)
    st = idx.stats()
    assert st["sa_sample_rate"] == rate and st["sa_resident"] == 1
    # arbitrary rows through the sampled array, including row 0 (the sentinel suffix) and the primary row
    rows = np.concatenate([[0, idx.primary, idx.length - 1], rng.integers(0, idx.length, 5000)]).astype(np.uint32)
    assert np.array_equal(idx.sa_lookup(rows, force_sampled=True), sa[rows])
    assert np.array_equal(idx.sa_lookup(rows), sa[rows])

    idx.sample_sa(rate, drop_sa=True)  # the full array leaves HBM; locate walks LF from now on
    st = idx.stats()
    assert st["sa_resident"] == 0 and idx.device_ptr("sa") == 0
    poff, pos = idx.locate(L, R)
    assert np.array_equal(poff, poff_e) and np.array_equal(pos, pos_e)
    # every row of the array, in order, equals the full array
    allrows = np.arange(idx.length, dtype=np.uint32)
    assert np.array_equal(idx.sa_lookup(allrows), sa)
    # ascending positions per pattern
    poff_s, pos_s = idx.locate(L, R, sorted=True)
    assert np.array_equal(poff_s, poff_e)
    exp = pos_e.copy()
    for q in range(len(L)):
        a, b = int(poff_e[q]), int(poff_e[q + 1])
        if b - a > 1:
            exp[a:b] = np.sort(exp[a:b])
    assert np.array_equal(pos_s, exp)
    idx.close()


def test_locate_without_any_suffix_array_fails(engine, oracle):
    codes = oracle.random_codes(5000, 4, seed=5)
    idx = engine.SuffixArrayIndex.build(codes[:-1], 5, drop_sa=True)
    with pytest.raises(engine.B200saError) as e:
        idx.locate([0], [3])
    assert e.value.code == 5  # NOT_BUILT
    with pytest.raises(engine.B200saError):
        idx.sample_sa(8)
    idx.close()


def test_sorted_locate_on_reference_texts(engine, oracle):
    """mississippi (the reference's own test string, bwt_test.c:16-38): sorted positions through the
    sampled array equal the known occurrences."""
    codes, sigma, table = oracle.remap(b"mississippi")
    idx = engine.SuffixArrayIndex.build(codes[:-1], sigma)
    idx.sample_sa(3, drop_sa=True)
    for raw, expected in ((b"ssi", [2, 5]), (b"i", [1, 4, 7, 10]), (b"mississippi", [0]), (b"x", None)):
        p = oracle.remap_pattern(table, raw)
        if p is None:
            assert expected is None
            continue
        L, R = idx.search_one(p)
        _, pos = idx.locate([L], [R], sorted=True)
        assert pos.tolist() == expected
    idx.close()
