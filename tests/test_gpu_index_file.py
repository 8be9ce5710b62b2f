"""GPU tests of the native index file (b200sa_save / b200sa_load, SURVEY 8f rank 1 for large n):
what is loaded answers exactly like what was saved -- same arrays, same (L, R), same positions."""
import numpy as np
import pytest

from test_gpu_parity import make_patterns

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,nsym,kw", [
    (300000, 4, dict(isa=True, lcp=True, bwt=True, textcmp=True, ktable=True)),
    (70001, 4, dict()),
    (50000, 20, dict(bwt=True, lcp=True)),
    (90000, 5, dict(isa=True, bwt=True, textcmp=True, ktable=True)),   # A C G N T: base-5 seed table, text comparison
    (60000, 20, dict(ktable=True)),
    (1, 4, dict(isa=True)),
])
def test_save_load_round_trip(engine, oracle, tmp_path, n, nsym, kw):
    rng = np.random.default_rng(n)
    codes = oracle.random_codes(n, nsym, seed=3 * n)
    sigma = nsym + 1
    idx = engine.SuffixArrayIndex.build(codes[:-1], sigma, **kw)
    idx.sample_sa(16)
    path = tmp_path / "index.b200sa"
    idx.save(path)
    back = engine.SuffixArrayIndex.load(path)
    assert (back.length, back.sigma, back.primary) == (idx.length, idx.sigma, idx.primary)
    sa = idx.sa()
    assert np.array_equal(back.sa(), sa) and np.array_equal(back.c_table(), idx.c_table())
    if kw.get("isa") or kw.get("textcmp"):
        assert np.array_equal(back.isa(), idx.isa())
    if kw.get("lcp"):
        assert np.array_equal(back.lcp(), idx.lcp())
    if kw.get("bwt"):
        assert np.array_equal(back.bwt(), idx.bwt())
    sa_ = back.stats()
    sb_ = idx.stats()
    assert sa_ == sb_
    pat, off = make_patterns(rng, codes, nsym, 3000, 1, 30)
    L, R = idx.search(pat, off)
    L2, R2 = back.search(pat, off)
    assert np.array_equal(L, L2) and np.array_equal(R, R2)
    assert all(np.array_equal(a, b) for a, b in zip(idx.locate(L, R), back.locate(L, R)))
    rows = rng.integers(0, idx.length, 2000).astype(np.uint32)
    assert np.array_equal(back.sa_lookup(rows, force_sampled=True), sa[rows])
    a, b = idx.approx_search(pat[:int(off[200])], off[:201], max_edits=1), back.approx_search(pat[:int(off[200])], off[:201], max_edits=1)
    assert a["cigars"] == b["cigars"] and np.array_equal(a["L"], b["L"]) and np.array_equal(a["offsets"], b["offsets"])
    idx.close()
    back.close()


def test_search_only_index_file_is_small(engine, oracle, tmp_path):
    """A replica for search + locate: O blocks, k-mer table and the sampled suffix array only."""
    n = 1 << 20
    codes = oracle.random_codes(n, 4, seed=11)
    idx = engine.SuffixArrayIndex.build(codes[:-1], 5, ktable=True)
    sa = idx.sa()
    idx.sample_sa(32, drop_sa=True)
    path = tmp_path / "search.b200sa"
    idx.save(path)
    # 0.5 B/row O blocks + 0.25 marks + 0.125 values + the k-mer table (k = 10 here: 8 MB), no 4 B/row SA
    assert path.stat().st_size < n * 1.0 + 8 * 4 ** 10 + 65536
    back = engine.SuffixArrayIndex.load(path)
    rng = np.random.default_rng(1)
    pat, off = make_patterns(rng, codes, 4, 5000, 12, 30)
    L, R = back.search(pat, off)
    poff, pos = back.locate(L, R)
    poff_e, pos_e = oracle.locate(sa, L, R)
    assert np.array_equal(poff, poff_e) and np.array_equal(pos, pos_e)
    with pytest.raises(engine.B200saError):
        back.sa()  # the full array is not in the file
    idx.close()
    back.close()


def test_load_rejects_bad_files(engine, tmp_path):
    bad = tmp_path / "bad.b200sa"
    bad.write_bytes(b"not an index" * 1000)
    with pytest.raises(engine.B200saError) as e:
        engine.SuffixArrayIndex.load(bad)
    assert e.value.code == 2
    with pytest.raises(engine.B200saError):
        engine.SuffixArrayIndex.load(tmp_path / "missing.b200sa")
    # a truncated file
    idx = engine.SuffixArrayIndex.build(np.array([1, 2, 3, 4] * 500, dtype=np.uint8), 5)
    good = tmp_path / "good.b200sa"
    idx.save(good)
    blob = good.read_bytes()
    (tmp_path / "cut.b200sa").write_bytes(blob[: len(blob) - 100])
    with pytest.raises(engine.B200saError):
        engine.SuffixArrayIndex.load(tmp_path / "cut.b200sa")
    idx.close()
