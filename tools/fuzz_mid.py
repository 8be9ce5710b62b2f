"""Mid-size randomised sweep (run under gpurun): texts of 2^20 .. 2^26 symbols with random repeat
structure, built with the thresholds the library picks by itself at these sizes (plus random knobs),
checked with the independent linear-time checker of stralg_b200/texts.py (SA), the definition of the
BWT (stralg/bwt.c:13-20) and the definition of the inverse.  Prints the first failure and exits
non-zero.
    python tools/fuzz_mid.py SECONDS SEED [MAXLOG2]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stralg_b200  # noqa: E402
from stralg_b200 import texts as T  # noqa: E402

lib = stralg_b200.load()
seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
maxlog = int(sys.argv[3]) if len(sys.argv) > 3 else 26
g = torch.Generator(device="cuda")
g.manual_seed(seed)
cpu = torch.Generator()
cpu.manual_seed(seed)
dev = torch.device("cuda:0")


def ri(lo, hi):
    return int(torch.randint(lo, hi, (1,), generator=cpu))


def rand_sym(n, nsym):
    return torch.randint(1, nsym + 1, (n,), generator=g, device=dev, dtype=torch.uint8 if nsym < 255 else torch.int16).to(torch.uint8)


def mutate(t, nsym, inv):
    if inv <= 0:
        return t
    hit = torch.randint(0, inv, (t.numel(),), generator=g, device=dev) == 0
    return torch.where(hit, rand_sym(t.numel(), nsym), t)


def make(kind, n, nsym):
    if kind == 0:   # uniform
        return rand_sym(n, nsym), "uniform"
    if kind == 1:   # copies of random segments of the text itself
        t = rand_sym(n, nsym)
        copies = ri(1, 4000)
        lmax = ri(2, max(3, min(n // 4, 200000)))
        for _ in range(min(copies, 300)):
            ln = ri(1, lmax)
            a, b = ri(0, n - ln), ri(0, n - ln)
            t[b:b + ln] = t[a:a + ln].clone()
        return t, f"copies lmax {lmax}"
    if kind == 2:   # a block tiled, with point mutations
        per = ri(1, max(2, n // 3))
        inv = [0, 0, 16, 64, 256, 4096, 65536][ri(0, 7)]
        base = rand_sym(per, nsym)
        t = base.repeat(n // per + 1)[:n].clone()
        return mutate(t, nsym, inv), f"tiled period {per} mutation 1/{inv}"
    if kind == 3:   # long runs
        cnt = n // 3 + 1
        vals = rand_sym(cnt, nsym)
        reps = torch.randint(1, ri(2, 40), (cnt,), generator=g, device=dev)
        return torch.repeat_interleave(vals, reps)[:n].clone(), "runs"
    if kind == 4:   # Fibonacci / Thue-Morse over two letters of the alphabet
        a, b = 1, min(2, nsym)
        if ri(0, 2):
            x, y = torch.tensor([a], dtype=torch.uint8, device=dev), torch.tensor([a, b], dtype=torch.uint8, device=dev)
            while y.numel() < n:
                x, y = y, torch.cat([y, x])
            return y[:n].clone(), "fibonacci"
        i = torch.arange(n, device=dev)
        par = torch.zeros(n, dtype=torch.int64, device=dev)
        for s in range(0, 32):
            par ^= (i >> s) & 1
        return torch.where(par == 0, torch.tensor(a, device=dev), torch.tensor(b, device=dev)).to(torch.uint8), "thue-morse"
    if kind == 5:   # one rare letter, in runs and sprinkled (N in DNA): the keys of a sparse alphabet must stay even
        t = rand_sym(n, max(1, nsym - 1))
        rare = nsym
        for _ in range(ri(0, 6)):
            ln = ri(1, max(2, n // 40))
            a0 = ri(0, max(1, n - ln))
            t[a0:a0 + ln] = rare
        idx = torch.randint(0, n, (max(1, n // ri(50, 100000)),), generator=g, device=dev)
        t[idx] = rare
        return t, "rare letter"
    # concatenation of differently structured parts
    parts, left, names = [], n, []
    while left > 0:
        ln = min(left, ri(1, max(2, n // 2)))
        p, nm = make(ri(0, 6), ln, nsym)
        parts.append(p[:ln])
        names.append(nm)
        left -= ln
    return torch.cat(parts)[:n].clone(), "mix(" + "; ".join(names) + ")"


KNOBS = ("B200SA_PIVOT_MIN", "B200SA_PIVOT_FORCE", "B200SA_PAIRS", "B200SA_ROUND0", "B200SA_SMALL_PATH",
         "B200SA_DENSE_FACTOR", "B200SA_CHAIN")
t_end = time.time() + seconds
cases = 0
while time.time() < t_end:
    n = ri(1 << 20, 1 << ri(21, maxlog + 1))
    nsym = [1, 2, 2, 3, 4, 4, 4, 5, 5, 6, 15, 16, 20, 100, 255][ri(0, 15)]
    text, name = make(ri(0, 7), n, nsym)
    n = text.numel()
    knobs = {}
    r = ri(0, 10)
    if r == 0:
        knobs["B200SA_PIVOT_FORCE"] = "1"
    elif r == 1:
        knobs["B200SA_PIVOT_MIN"] = str(1 << ri(1, 22))
    elif r == 2:
        knobs["B200SA_PAIRS"] = str(ri(0, 3))
    elif r == 3:
        knobs["B200SA_ROUND0"] = "lsd"
    elif r == 4:
        knobs["B200SA_SMALL_PATH"] = "0"
    for k in KNOBS:
        os.environ.pop(k, None)
    os.environ.update(knobs)
    tag = {"case": cases, "seed": seed, "n": n, "nsym": nsym, "text": name[:200], "knobs": knobs}
    t0 = time.time()
    want_isa = bool(ri(0, 2))
    idx = stralg_b200.SuffixArrayIndex.build(text, nsym + 1, isa=want_isa, bwt=True, occ=True,
                                             textcmp=bool(ri(0, 2)), ktable=bool(ri(0, 2)))
    torch.cuda.synchronize()
    tag["build_ms"] = round((time.time() - t0) * 1e3, 1)
    st = idx.stats()
    tag.update({k: st[k] for k in ("rounds", "round0_mode", "pivot_rounds", "pair_placed", "chain_rounds")})
    sa = T.device_view(idx.device_ptr("sa"), n + 1, 4)
    tz = torch.cat([text, torch.zeros(1, dtype=torch.uint8, device=dev)])
    ok, why = T.check_suffix_array(tz, sa, n)
    if not ok:
        print("FAIL SA", why, json.dumps(tag), flush=True)
        torch.save(text.cpu(), "gpurun_out/fuzz_mid_fail.pt")
        sys.exit(1)
    s = sa.long() & 0xFFFFFFFF
    bwt = T.device_view(idx.device_ptr("bwt"), n + 1, 1)
    exp = torch.where(s == 0, torch.zeros_like(s, dtype=torch.uint8), tz[torch.clamp(s - 1, min=0)])
    if not bool((bwt == exp).all()):
        print("FAIL BWT", json.dumps(tag), flush=True)
        torch.save(text.cpu(), "gpurun_out/fuzz_mid_fail.pt")
        sys.exit(1)
    if want_isa:
        isa = T.device_view(idx.device_ptr("isa"), n + 1, 4).long() & 0xFFFFFFFF
        if not bool((s[isa] == torch.arange(n + 1, device=dev)).all()):
            print("FAIL ISA", json.dumps(tag), flush=True)
            sys.exit(1)
        del isa
    # patterns cut from the text are found, and the interval's first row points at an occurrence
    m = ri(1, 60)
    npat = 2000
    starts = torch.randint(0, max(1, n - m), (npat,), generator=g, device=dev)
    pat = tz[(starts[:, None] + torch.arange(m, device=dev)[None, :]).clamp(max=n)]
    keep = (pat != 0).all(dim=1)
    pat = pat[keep].contiguous()
    if pat.numel():
        dL = torch.empty(pat.shape[0], dtype=torch.int32, device=dev)
        dR = torch.empty_like(dL)
        idx.search_device(pat, None, m, pat.shape[0], dL, dR)
        torch.cuda.synchronize()
        L, R = dL.long() & 0xFFFFFFFF, dR.long() & 0xFFFFFFFF
        if not bool((L < R).all()):
            print("FAIL search: a pattern cut from the text was not found", json.dumps(tag), flush=True)
            sys.exit(1)
        first = s[L]
        got = tz[(first[:, None] + torch.arange(m, device=dev)[None, :]).clamp(max=n)]
        if not bool((got == pat).all()):
            print("FAIL search: row L does not point at an occurrence", json.dumps(tag), flush=True)
            sys.exit(1)
        last = s[R - 1]
        got = tz[(last[:, None] + torch.arange(m, device=dev)[None, :]).clamp(max=n)]
        if not bool((got == pat).all()):
            print("FAIL search: row R-1 does not point at an occurrence", json.dumps(tag), flush=True)
            sys.exit(1)
    print(json.dumps(tag), flush=True)
    idx.close()
    del text, tz, sa, s, bwt, exp
    torch.cuda.empty_cache()
    cases += 1
print(f"ok: {cases} cases in {seconds:.0f} s (seed {seed})")
