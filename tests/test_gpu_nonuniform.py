"""Non-uniform texts (BASELINE.json configs[4]; SURVEY 8(d) C3 repeat-rich / C5).

The reference's constructors are input-insensitive (stralg/sa_is.c:340-441 is linear on any text;
performance/suffix_array_construction.c:81-184 times "Equal" strings next to random ones), so the
bucketed round 0 must not depend on a flat k-mer spectrum:

* a bucket too large for one SM (poly-A runs, tandem / interspersed repeats) is emitted as a shallow
  group and ordered by the doubling rounds (round0_msd.cu: OverArgs),
* a skewed spectrum (real DNA) adds a partition level,
* only texts whose oversize buckets hold more than 1/8 of the suffixes take the LSD path.

Small sizes: SA / ISA / LCP / BWT bit-exact against the reference's SA-IS (oracle/_ref) or the
restatement.  Large sizes (256 Mi repeat-rich, 2^30 a^n): the suffix-array checker of
stralg_b200/texts.py, independent of the product kernels.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _skewed(rng, n, p):
    return rng.choice(np.arange(1, len(p) + 1), size=n, p=p).astype(np.uint8)


def nonuniform_texts():
    from stralg_b200 import texts as T
    rng = np.random.default_rng(4242)
    out = []
    base = rng.integers(1, 5, 1 << 20).astype(np.uint8)
    # poly-A in the middle and at the very end: the A^k bucket is oversize and holds the short suffixes
    t = base.copy()
    t[400000:420000] = 1
    t[-5000:] = 1
    out.append(("polyA_mid_end_1M", t, 5))
    # the same with the run ending one symbol before the end, and a (CA)^k tandem repeat
    t = base.copy()
    t[-9000:-1] = 1
    t[100000:130000] = np.tile([2, 1], 15000)
    out.append(("polyA_tandemCA_1M", t, 5))
    # three copies of a 60 k block and twenty of a 4 k block (deep repeats, few copies)
    t = base.copy()
    blk = t[5000:65000].copy()
    t[300000:360000] = blk
    t[700000:760000] = blk
    small = t[900000:904000].copy()
    for k in range(20):
        t[20000 * k + 1000: 20000 * k + 5000] = small
    out.append(("block_copies_1M", t, 5))
    # skewed symbol distribution (70 % A): the planned average bucket says little, a level is added
    out.append(("skewed_70A_1M", _skewed(rng, 1 << 20, [0.7, 0.1, 0.1, 0.1]), 5))
    out.append(("skewed_97A_300k", _skewed(rng, 300000, [0.97, 0.01, 0.01, 0.01]), 5))
    # real DNA: the genome sample, once and tiled with 1.6 % point mutations
    hg = T.hg38_base()
    out.append(("hg38_sample_500k", hg, 5))
    tiled = np.tile(hg[:200000], 6)
    mut = rng.random(len(tiled)) < 1 / 64
    tiled[mut] = (tiled[mut] - 1 + rng.integers(1, 4, int(mut.sum()))) % 4 + 1
    out.append(("hg38_tiled_mut_1200k", tiled.astype(np.uint8), 5))
    # other alphabets: binary with long runs, 16 letters with one dominant letter
    runs = np.repeat(rng.integers(1, 3, 40000), rng.integers(1, 60, 40000))[:600000].astype(np.uint8)
    out.append(("binary_runs_600k", runs, 3))
    out.append(("sym16_skewed_500k", _skewed(rng, 500000, [0.85] + [0.01] * 15), 17))
    out.append(("byte_skewed_400k", _skewed(rng, 400000, [0.9] + [0.1 / 254] * 254), 256))
    # exact copies of segments, two of each (pairs of equal keys; pair_runs_kernel): plain, overlapping an earlier copy,
    # tandem (distance shorter than the segment), one ending at the end of the text, one longer than a tile
    t = base.copy()
    for k in range(60):
        ln = int(rng.integers(300, 6000))
        src = int(rng.integers(0, len(t) - ln))
        dst = int(rng.integers(0, len(t) - ln))
        t[dst:dst + ln] = t[src:src + ln].copy()
    t[500000:503000] = t[500100:503100].copy()
    t[-4000:] = t[200000:204000].copy()
    t[40000:70000] = t[800000:830000].copy()
    out.append(("pair_copies_1M", t, 5))
    # the same on a binary alphabet (64 symbols per compared window) and with a long run inside the copies
    t2 = rng.integers(1, 3, 700000).astype(np.uint8)
    for k in range(30):
        ln = int(rng.integers(500, 9000))
        src = int(rng.integers(0, len(t2) - ln))
        dst = int(rng.integers(0, len(t2) - ln))
        t2[dst:dst + ln] = t2[src:src + ln].copy()
    t2[100000:100700] = 1
    t2[400000:400700] = 1
    out.append(("pair_copies_binary_700k", t2, 3))
    # alphabets that do not fill their symbol width (dense initial keys, round0_msd.cuh DenseKey): DNA with N
    # (A C G N T, N in long runs and sprinkled, one run at the very end), amino acids with copied segments,
    # three letters with a period and a run of the smallest letter at the end
    t = rng.choice(np.array([1, 2, 3, 5], dtype=np.uint8), 1 << 20)
    for k in range(5):  # (about 5 % of the text, as in a real assembly: more than 1/8 would take the LSD path)
        a = int(rng.integers(0, len(t) - 60000))
        t[a:a + int(rng.integers(500, 20000))] = 4
    t[rng.integers(0, len(t), 300)] = 4
    t[-3000:] = 4
    out.append(("dna_n_runs_1M", t, 6))
    t = rng.integers(1, 21, 800000).astype(np.uint8)
    for k in range(40):
        ln = int(rng.integers(200, 20000))
        src = int(rng.integers(0, len(t) - ln))
        dst = int(rng.integers(0, len(t) - ln))
        t[dst:dst + ln] = t[src:src + ln].copy()
    t[300000:330000] = 7
    out.append(("amino_copies_800k", t, 21))
    t = rng.integers(1, 4, 600000).astype(np.uint8)
    t[100000:160000] = np.tile(rng.integers(1, 4, 37).astype(np.uint8), 60000 // 37 + 1)[:60000]
    t[-2000:] = 1
    out.append(("ternary_period_600k", t, 4))
    return out


TEXTS = None


def _texts():
    global TEXTS
    if TEXTS is None:
        TEXTS = {t[0]: t for t in nonuniform_texts()}
    return TEXTS


NAMES = ["polyA_mid_end_1M", "polyA_tandemCA_1M", "block_copies_1M", "skewed_70A_1M", "skewed_97A_300k",
         "hg38_sample_500k", "hg38_tiled_mut_1200k", "binary_runs_600k", "sym16_skewed_500k", "byte_skewed_400k",
         "pair_copies_1M", "pair_copies_binary_700k", "dna_n_runs_1M", "amino_copies_800k", "ternary_period_600k"]

SPARSE_ALPHABETS = ("dna_n_runs_1M", "amino_copies_800k", "ternary_period_600k", "sym16_skewed_500k")

# default plan; never add a level (everything oversize becomes a shallow group, however much it is);
# always add a level when one bucket is oversize; tiny buckets (many segments per tile) with shallow groups
MODES = {
    "default": {},
    "shallow_only": {"B200SA_MSD_MORE_FRAC": "1", "B200SA_MSD_OVER_FRAC": "1"},
    "more_levels": {"B200SA_MSD_MORE_FRAC": "100000000", "B200SA_MSD_OVER_FRAC": "1"},
    "avg64_shallow": {"B200SA_MSD_AVG": "64", "B200SA_MSD_MORE_FRAC": "1", "B200SA_MSD_OVER_FRAC": "1"},
    # chain offsets (sa_build.cu chain_flags_kernel) on every round however small the active set / never
    "chain_always": {"B200SA_CHAIN_MIN_FRAC": "100000000", "B200SA_CHAIN_USE_FRAC": "100000000"},
    "chain_always_shallow": {"B200SA_CHAIN_MIN_FRAC": "100000000", "B200SA_CHAIN_USE_FRAC": "100000000", "B200SA_MSD_MORE_FRAC": "1", "B200SA_MSD_OVER_FRAC": "1"},
    "chain_off": {"B200SA_CHAIN": "0"},
    # every group goes through the radix sort (no in-tile ordering of small groups, no text tie-break)
    "all_radix": {"B200SA_SMALL_PATH": "0", "B200SA_NO_EXT_TIEBREAK": "1"},
    # groups split around a pivot key in every round / by the path's own heuristics on lists of any size
    "pivot_force": {"B200SA_PIVOT_MIN": "2", "B200SA_PIVOT_FORCE": "1"},
    "pivot_auto": {"B200SA_PIVOT_MIN": "2", "B200SA_SMALL_PATH": "0"},
    # groups of two decided by one text comparison per repeat, on lists of any size; with and without the tie-break
    # by the next 64 bits of text in front of it; never
    "pairs_force": {"B200SA_PAIRS": "2"},
    "pairs_force_noext": {"B200SA_PAIRS": "2", "B200SA_NO_EXT_TIEBREAK": "1"},
    "pairs_off": {"B200SA_PAIRS": "0"},
    # initial keys as base-(letters) numbers on every alphabet / never; with shallow groups only
    "dense_keys_on": {"B200SA_DENSE_KEYS": "1"},
    "dense_keys_off": {"B200SA_DENSE_KEYS": "0"},
    "dense_keys_shallow": {"B200SA_DENSE_KEYS": "1", "B200SA_MSD_MORE_FRAC": "1", "B200SA_MSD_OVER_FRAC": "1"},
}


@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("name", NAMES)
def test_nonuniform_tables_match_oracle(engine, oracle, ref, name, mode, monkeypatch):
    for k, v in MODES[mode].items():
        monkeypatch.setenv(k, v)
    _, sym, sigma = _texts()[name]
    codes = np.concatenate([np.asarray(sym, dtype=np.uint8), np.zeros(1, np.uint8)])
    idx = engine.SuffixArrayIndex.build(codes[:-1], sigma, isa=True, lcp=True, bwt=True, occ=True)
    st = idx.stats()
    if "B200SA_MSD_OVER_FRAC" in MODES[mode] and mode != "dense_keys_shallow" and name not in SPARSE_ALPHABETS:
        # OVER_FRAC = 1: nothing may fall back to the LSD path (with dense keys an oversize bucket that shares too few
        # symbols still does)
        assert st["round0_mode"] == 1, (name, mode, st)
    sa_exp = ref.sa(codes, sigma, "sa_is") if ref is not None else oracle.sa(codes)
    sa = idx.sa()
    assert np.array_equal(sa, sa_exp), f"{name}/{mode}: SA differs at {np.nonzero(sa != sa_exp)[0][:5]}, stats {st}"
    assert np.array_equal(idx.isa(), oracle.inverse(sa_exp))
    lcp_exp = oracle.lcp(codes, sa_exp)
    lcp = idx.lcp()
    assert np.array_equal(lcp, lcp_exp), f"{name}/{mode}: LCP differs at {np.nonzero(lcp != lcp_exp)[0][:5]}"
    bwt_exp = oracle.bwt(codes, sa_exp)
    assert np.array_equal(idx.bwt(), bwt_exp), f"{name}/{mode}: BWT differs"
    assert idx.primary == int(np.nonzero(sa_exp == 0)[0][0])
    ck = oracle.o_checkpoints(bwt_exp, sigma, 64)
    rng = np.random.default_rng(3)
    qa = rng.integers(0, sigma, 3000).astype(np.uint8)
    qi = rng.integers(0, len(sa) + 1, 3000).astype(np.uint32)
    assert np.array_equal(idx.occ(qa, qi), oracle.o_probe(bwt_exp, ck, sigma, 64, qa, qi))
    print(f"[nonuniform] {name}/{mode}: {st}")
    idx.close()


def test_shallow_groups_are_taken(engine):
    """The poly-A text must really go through the shallow-group path (no silent LSD fallback), and the
    skewed one must really get a level added."""
    _, sym, sigma = _texts()["polyA_mid_end_1M"]
    idx = engine.SuffixArrayIndex.build(np.asarray(sym, dtype=np.uint8), sigma)
    st = idx.stats()
    assert st["round0_mode"] == 1 and st["shallow_buckets"] >= 1, st
    idx.close()
    _, sym, sigma = _texts()["skewed_70A_1M"]
    idx = engine.SuffixArrayIndex.build(np.asarray(sym, dtype=np.uint8), sigma)
    st2 = idx.stats()
    assert st2["round0_mode"] == 1 and st2["passes0"] >= 2, st2
    idx.close()


def test_dense_keys_are_taken(engine):
    """Sparse alphabets must really get dense initial keys and stay on the bucketed round 0 -- the DNA + N text
    with its N runs as shallow groups -- and a full alphabet must not."""
    for name, want_shallow in (("dna_n_runs_1M", True), ("amino_copies_800k", False), ("ternary_period_600k", False)):
        _, sym, sigma = _texts()[name]
        idx = engine.SuffixArrayIndex.build(np.asarray(sym, dtype=np.uint8), sigma)
        st = idx.stats()
        assert st["round0_mode"] == 1 and st["dense_keys"] == sigma - 1, (name, st)
        if want_shallow:
            assert st["shallow_buckets"] >= 1, (name, st)
        idx.close()
    _, sym, sigma = _texts()["pair_copies_1M"]
    idx = engine.SuffixArrayIndex.build(np.asarray(sym, dtype=np.uint8), sigma)
    assert idx.stats()["dense_keys"] == 0
    idx.close()


def test_two_devices_one_process(engine):
    """the opt-in for > 48 KB of dynamic shared memory is per device; a process that
    builds on a second GPU must not fail.  Runs only where two devices are visible."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("one visible device")
    rng = np.random.default_rng(5)
    sym = rng.integers(1, 5, 300000).astype(np.uint8)
    a = engine.SuffixArrayIndex.build(sym, 5, device=0)
    b = engine.SuffixArrayIndex.build(sym, 5, device=1)
    assert np.array_equal(a.sa(), b.sa())
    a.close()
    b.close()


def test_sharded_search_behind_the_c_abi(engine, oracle):
    """b200sa_replicate + b200sa_search_sharded_packed (one process, one replica per
    device, contiguous read shards, (L, R) stored into the first replica's HBM by the search kernels):
    equal to the single-device search.  With one visible device the single-replica path is checked."""
    import torch
    ndev = min(torch.cuda.device_count(), 4)
    n, m, npat = 1 << 20, 100, 200003
    codes = oracle.random_codes(n, 4, seed=5)
    idx = engine.SuffixArrayIndex.build(codes[:-1], 5, textcmp=True, ktable=True)
    rng = np.random.default_rng(9)
    starts = rng.integers(0, n - m, npat)
    reads = np.stack([codes[s:s + m] for s in starts]).astype(np.uint8)
    reads[::7] = rng.integers(1, 5, (len(reads[::7]), m))
    packed = engine.pack_reads(reads.reshape(-1), m)
    L, R = idx.search_packed(packed, m, npat)
    reps = [idx] + [idx.replicate(d) for d in range(1, ndev)]
    for k in range(1, ndev):
        assert np.array_equal(reps[k].sa_lookup(np.arange(0, n, 997, dtype=np.uint32)),
                              idx.sa_lookup(np.arange(0, n, 997, dtype=np.uint32)))
    for use in range(1, ndev + 1):
        Ls, Rs = engine.search_sharded_packed(reps[:use], packed, m, npat)
        assert np.array_equal(Ls, L) and np.array_equal(Rs, R), use
    print(f"[sharded-c] {ndev} device(s): (L, R) of {npat} reads equal to the single-device search")
    for r in reps[1:]:
        r.close()
    idx.close()


# ---- large sizes: properties -----------------------------------------------------------------------
def _build_and_check(engine, text, n, sigma, label, **kw):
    import time
    import torch
    from stralg_b200 import texts as T
    lib = engine.load()
    torch.cuda.synchronize()
    t0 = time.time()
    idx = engine.SuffixArrayIndex.build(text[:n], sigma, occ=False, **kw)
    torch.cuda.synchronize()
    dt = time.time() - t0
    st = idx.stats()
    lib.b200sa_release_workspace(0)
    sa = T.device_view(idx.device_ptr("sa"), n + 1, 4)
    ok, why = T.check_suffix_array(text, sa, n)
    print(f"[nonuniform-large] {label}: n = {n}, build {dt * 1e3:.0f} ms ({n / dt / 1e6:.0f} Mchar/s), stats = {st}")
    assert ok, (label, why, st)
    idx.close()
    return st, dt


def test_repeat_rich_256M(engine):
    """SURVEY 8(d) C3 repeat-rich variant at 256 Mi: copies of random 300-6 000 bp segments (10 % of the text
    duplicated).  Must stay on the bucketed round 0."""
    import torch
    from stralg_b200 import texts as T
    lib = engine.load()
    n = int(float(os.environ.get("B200SA_NONUNIFORM_N", 1 << 28)))
    text = T.random_codes(lib, n, 4, 31337)
    info = T.add_repeats(text, n)
    torch.cuda.synchronize()
    st, _ = _build_and_check(engine, text, n, 5, f"repeat-rich {info}")
    assert st["round0_mode"] == 1, st


@pytest.mark.parametrize("kind", ["acgt4", "period1000", "fib"])
def test_periodic_pivot_path_64M(engine, kind):
    """Config 5 (C5c-e) at 2^26: the doubling rounds of periodic texts must run on the pivot path (the majority of
    every group keeps its place, only the minority is sorted) and yield the suffix array the checker accepts."""
    from stralg_b200 import texts as T
    n = int(float(os.environ.get("B200SA_PERIODIC_N", 1 << 26)))
    text, sigma = T.stress_text(engine.load(), kind, n)
    st, _ = _build_and_check(engine, text, n, sigma, kind)
    assert st["pivot_rounds"] >= 10 and st["pivot_elems"] > 4 * n, st


def test_unary_2pow30(engine):
    """Config 5 (C5b): a^n at n = 2^30, the worst-case doubling depth."""
    import torch
    from stralg_b200 import texts as T
    free, _ = torch.cuda.mem_get_info()
    n = int(float(os.environ.get("B200SA_UNARY_N", (1 << 30) if free > 100e9 else (1 << 26))))
    text, sigma = T.stress_text(engine.load(), "unary", n)
    st, _ = _build_and_check(engine, text, n, sigma, "a^n")
    assert st["rounds"] >= 20 or n < (1 << 30)
