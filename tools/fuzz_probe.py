"""Randomised parity sweep (run under gpurun): many small texts of random length, alphabet and
repeat structure; SA / ISA / LCP / BWT / C, exact search + locate (full and sampled SA, sorted),
approximate search (d = 0..2, with and without the reverse index) against the oracle.  Prints the
first mismatch and exits non-zero."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import stralg_b200
from _oracle import Oracle

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
o = Oracle()
t_end = time.time() + seconds
cases = 0
while time.time() < t_end:
    kind = int(rng.integers(0, 5))
    n = int(rng.integers(1, 6000)) if kind != 4 else int(rng.integers(50000, 300000))
    nsym = int(rng.choice([1, 2, 3, 4, 4, 4, 5, 16, 17, 100, 255]))
    if kind == 0 or kind == 4:
        sym = rng.integers(1, nsym + 1, n)
    elif kind == 1:  # periodic
        per = int(rng.integers(1, 40))
        sym = np.tile(rng.integers(1, nsym + 1, per), n // per + 1)[:n]
    elif kind == 2:  # long runs
        sym = np.repeat(rng.integers(1, nsym + 1, n // 7 + 1), rng.integers(1, 15, n // 7 + 1))[:n]
        n = len(sym)
    else:  # repeats of random segments
        seg = rng.integers(1, nsym + 1, max(1, n // 5))
        sym = np.concatenate([seg, rng.integers(1, nsym + 1, max(1, n // 3)), seg, seg[: len(seg) // 2]])[:n]
        n = len(sym)
    codes = np.concatenate([sym.astype(np.uint8), np.zeros(1, np.uint8)])
    sigma = nsym + 1
    # the paths of the doubling rounds that real sizes choose by themselves, forced at random on these small texts:
    # the pivot path (sa_build.cu pivot_classify_kernel), the pair path (pair_runs_kernel), the LSD round 0
    knobs = {}
    if rng.integers(0, 2):
        knobs["B200SA_PIVOT_MIN"] = "2"
        if rng.integers(0, 2):
            knobs["B200SA_PIVOT_FORCE"] = "1"
    if rng.integers(0, 2):
        knobs["B200SA_PAIRS"] = "2"
    if rng.integers(0, 4) == 0:
        knobs["B200SA_ROUND0"] = "lsd"
    if rng.integers(0, 3) == 0:
        knobs["B200SA_SMALL_PATH"] = "0"
    for k in ("B200SA_PIVOT_MIN", "B200SA_PIVOT_FORCE", "B200SA_PAIRS", "B200SA_ROUND0", "B200SA_SMALL_PATH"):
        os.environ.pop(k, None)
    os.environ.update(knobs)
    tag = f"case {cases} seed {seed} kind {kind} n {n} sigma {sigma} knobs {knobs}"
    idx = stralg_b200.SuffixArrayIndex.build(codes[:-1], sigma, isa=True, lcp=True, bwt=True, occ=True,
                                             textcmp=bool(rng.integers(0, 2)), ktable=bool(rng.integers(0, 2)))
    sa_e = o.sa(codes)
    sa = idx.sa()
    assert np.array_equal(sa, sa_e), tag + " SA"
    isa_e = o.inverse(sa_e)
    assert np.array_equal(idx.isa(), isa_e), tag + " ISA"
    assert np.array_equal(idx.lcp(), o.lcp(codes, sa_e, isa_e)), tag + " LCP"
    bwt_e = o.bwt(codes, sa_e)
    assert np.array_equal(idx.bwt(), bwt_e), tag + " BWT"
    c_e = o.c_table(codes, sigma)
    assert np.array_equal(idx.c_table(), c_e), tag + " C"
    # patterns
    npat = 300
    pats = []
    for k in range(npat):
        m = int(rng.integers(1, 25))
        if k % 2 and n > m:
            s0 = int(rng.integers(0, n - m))
            p = codes[s0:s0 + m].copy()
            if k % 4 == 1:
                p[int(rng.integers(0, m))] = 1 + int(rng.integers(0, nsym))
        else:
            p = rng.integers(1, nsym + 1, m).astype(np.uint8)
        pats.append(p)
    off = np.zeros(npat + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(p) for p in pats])
    pat = np.concatenate(pats).astype(np.uint8)
    L, R = idx.search(pat, off)
    if n <= 6000:
        otab = o.o_table(bwt_e, sigma)
        Le, Re = o.search_dense(c_e, otab, len(codes), pat, off)
    else:
        ck = o.o_checkpoints(bwt_e, sigma, 64)
        Le, Re = o.search_ck(c_e, bwt_e, ck, 64, pat, off)
    assert np.array_equal(L, Le) and np.array_equal(R, Re), tag + " (L,R)"
    # one pattern per call (the resident kernel of a DNA index, the launch path otherwise)
    for q in range(0, npat, 37):
        a, b = int(off[q]), int(off[q + 1])
        L1, R1 = idx.search(np.concatenate([pat[a:b], np.zeros(8, np.uint8)]), np.array([0, b - a], dtype=np.uint64))
        assert (int(L1[0]), int(R1[0])) == (int(Le[q]), int(Re[q])), tag + f" one pattern q={q}"
    poff_e, pos_e = o.locate(sa_e, Le, Re)
    poff, pos = idx.locate(L, R)
    assert np.array_equal(poff, poff_e) and np.array_equal(pos, pos_e), tag + " locate"
    rate = int(rng.integers(1, 70))
    idx.sample_sa(rate, drop_sa=False)
    rows = rng.integers(0, len(codes), 500).astype(np.uint32)
    assert np.array_equal(idx.sa_lookup(rows, force_sampled=True), sa_e[rows]), tag + f" sampled SA rate {rate}"
    _, pos_s = idx.locate(L, R, sorted=True)
    exp = pos_e.copy()
    for q in range(npat):
        a, b = int(poff_e[q]), int(poff_e[q + 1])
        if b - a > 1:
            exp[a:b] = np.sort(exp[a:b])
    assert np.array_equal(pos_s, exp), tag + " sorted locate"
    # approximate search on the small cases (the oracle walks dense tables)
    if n <= 3000 and nsym <= 17:
        otab = o.o_table(bwt_e, sigma)
        rcodes = np.concatenate([codes[:-1][::-1], np.zeros(1, np.uint8)])
        rsa = o.sa(rcodes)
        ro = o.o_table(o.bwt(rcodes, rsa), sigma)
        use_rev = bool(rng.integers(0, 2))
        rev = stralg_b200.SuffixArrayIndex.build(rcodes[:-1], sigma, drop_sa=True) if use_rev else None
        d = int(rng.integers(0, 3)) if nsym <= 5 else int(rng.integers(0, 2))
        sub = 60
        res = idx.approx_search(pat[: int(off[sub])], off[: sub + 1], max_edits=d, rev=rev)
        h = 0
        for q in range(sub):
            La, Ra, ml, cig = o.approx(c_e, otab, ro if use_rev else None, len(codes), pats[q], d)
            a, b = int(res["offsets"][q]), int(res["offsets"][q + 1])
            assert np.array_equal(res["L"][a:b], La) and np.array_equal(res["R"][a:b], Ra), tag + f" approx d={d} q={q} (L,R)"
            assert np.array_equal(res["match_length"][a:b], ml) and res["cigars"][a:b] == cig, tag + f" approx d={d} q={q} cigar"
            h += b - a
        if rev is not None:
            rev.close()
    idx.close()
    cases += 1
print(f"fuzz ok: {cases} cases in {seconds:.0f} s (seed {seed})")
