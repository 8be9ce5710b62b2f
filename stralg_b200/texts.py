"""Deterministic synthetic texts on the device (bench.py, the GPU tests, tools/): remapped codes
1..sigma-1 followed by the sentinel 0, as torch uint8 tensors of n + 1 bytes.

SURVEY.md 8(d) names them: C3 random ACGT and its repeat-rich variant (copies of random
300-6 000 bp segments), C5a random bytes, C5b a^n, C5c (ACGT)^k, C5d a period-1000 random block,
C5e the Fibonacci string; the reference's own harness times "Equal" (unary) strings next to random
ones (performance/suffix_array_construction.c:81-184).  `hg38_like` tiles the reference's 500 kbp
human-genome sample (tools/readmappers/data/genomes/hg38-10000.fa, committed 2-bit packed as
tests/golden/hg38_10000.2bit.npy) with point mutations, so that the text has the k-mer spectrum of
real DNA (poly-A, Alu, tandem repeats) instead of a uniform one.

Only input generation lives here -- plain torch indexing, no product kernel and no oracle.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

SEED = 88172645463325252


def _mix(x, torch):
    """splitmix64-style hash of an int64 tensor (wrap-around arithmetic)."""
    x = (x ^ (x >> 30)) * -4658895280553007687  # 0xBF58476D1CE4E5B9
    x = (x ^ (x >> 27)) * -7723592293110705685  # 0x94D049BB133111EB
    return x ^ (x >> 31)


def random_codes(lib, n, nsym, seed, device=0, stream=None):
    import torch
    text = torch.empty(n + 1, dtype=torch.uint8, device=torch.device("cuda", device))
    rc = lib.b200sa_synth_codes(C.c_void_p(text.data_ptr()), n, nsym, seed, device,
                                C.c_void_p(stream) if stream else None)
    assert rc == 0
    return text


def add_repeats(text, n, copies=None, lmin=300, lmax=6000, seed=12345):
    """The repeat-rich variant of SURVEY 8(d) C3, in place: `copies` (default n / 30 000, i.e. 10^5 at
    3 Gbp) segments of lmin..lmax symbols, each read from a uniformly random place of the ORIGINAL text
    and written into its own slot of n / copies symbols (destinations are disjoint, so the result does
    not depend on the order of the copies).  About 10 % of the text ends up duplicated; a source that
    overlaps another copy's destination or source gives runs that occur three or more times."""
    import torch
    dev = text.device
    if copies is None:
        copies = max(1, n // 30000)
    slot = n // copies
    lmax = min(lmax, slot - 1)
    lmin = min(lmin, lmax)
    g = torch.Generator(device="cpu").manual_seed(seed)
    L = torch.randint(lmin, lmax + 1, (copies,), generator=g, dtype=torch.int64)
    src = (torch.rand(copies, generator=g, dtype=torch.float64) * (n - L).double()).long()
    dst = torch.arange(copies, dtype=torch.int64) * slot + (torch.rand(copies, generator=g, dtype=torch.float64)
                                                           * (slot - L).double()).long()
    step = 2000  # copies per batch: bounds the index tensors (<= 12 M entries)
    vals = []
    for lo in range(0, copies, step):
        l, s = L[lo:lo + step].to(dev), src[lo:lo + step].to(dev)
        off = torch.cumsum(l, 0) - l
        tot = int(l.sum())
        seg = torch.repeat_interleave(torch.arange(len(l), device=dev), l)
        k = torch.arange(tot, device=dev) - off[seg]
        vals.append(text[s[seg] + k].clone())  # every source is read before any destination is written
        del seg, k
    for i, lo in enumerate(range(0, copies, step)):
        l, d = L[lo:lo + step].to(dev), dst[lo:lo + step].to(dev)
        off = torch.cumsum(l, 0) - l
        tot = int(l.sum())
        seg = torch.repeat_interleave(torch.arange(len(l), device=dev), l)
        k = torch.arange(tot, device=dev) - off[seg]
        text[d[seg] + k] = vals[i]
        del seg, k
    return {"copies": int(copies), "copied_symbols": int(L.sum()), "lmin": int(lmin), "lmax": int(lmax)}


def hg38_base(path=None):
    """The 499 950 bases of the sample as codes 1..4 (numpy)."""
    if path is None:
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                            "hg38_10000.2bit.npy")
    raw = np.load(path)
    nb = int(np.frombuffer(raw[:4].tobytes(), dtype=np.uint32)[0])
    p = raw[4:]
    sym = np.stack([(p >> (2 * k)) & 3 for k in range(4)], axis=1).reshape(-1)[:nb]
    return (sym + 1).astype(np.uint8)


def hg38_like(n, device=0, mut_inv=64, seed=7, base=None):
    """Tiles the genome sample up to n symbols; position i is substituted by one of the three other
    letters when hash(i) % mut_inv == 0 (mut_inv = 64: 1.6 % divergence between copies)."""
    import torch
    dev = torch.device("cuda", device)
    b = torch.from_numpy(hg38_base() if base is None else base).to(dev)
    text = torch.zeros(n + 1, dtype=torch.uint8, device=dev)
    chunk = 1 << 27
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        i = torch.arange(lo, hi, device=dev, dtype=torch.int64)
        v = b[i % len(b)]
        h = _mix(i + seed * 1000003, torch)
        mut = (h % mut_inv) == 0
        sh = ((h >> 20) % 3 + 1).to(torch.uint8)
        v = torch.where(mut, ((v - 1 + sh) % 4) + 1, v)
        text[lo:hi] = v
        del i, v, h, mut, sh
    return text


def dna_with_n(lib, n, device=0, seed=3):
    """Random ACGT with N the way an assembly has it: codes A C G N T = 1 2 3 4 5 (sigma 6), about 5 % N in
    n / 10^7 runs of 10^3 .. 10^6 symbols (gaps, centromeres) plus one single N per 10^5 symbols.  Returns the text
    (n + 1 codes, sentinel last) and the fraction of N."""
    import torch
    t = random_codes(lib, n, 4, SEED, device)
    t[:n][t[:n] == 4] = 5
    g = torch.Generator().manual_seed(seed)
    for _ in range(max(1, n // 10_000_000)):
        ln = min(int(torch.randint(1000, 1_000_000, (1,), generator=g)), max(1, n // 20))
        a = int(torch.randint(0, max(1, n - ln), (1,), generator=g))
        t[a:a + ln] = 4
    idx = torch.randint(0, n, (max(1, n // 100000),), generator=g).to(t.device)
    t[idx] = 4
    return t, float((t[:n] == 4).float().mean())


STRESS_KINDS = ("byte", "unary", "acgt4", "period1000", "fib")


def stress_text(lib, kind, n, device=0):
    """Config 5: returns (text, sigma)."""
    import torch
    dev = torch.device("cuda", device)
    if kind == "byte":
        return random_codes(lib, n, 255, 5, device), 256
    text = torch.zeros(n + 1, dtype=torch.uint8, device=dev)
    if kind == "unary":
        text[:n] = 1
        return text, 2
    if kind == "acgt4":
        text[:n] = torch.tensor([1, 2, 3, 4], dtype=torch.uint8, device=dev).repeat(n // 4 + 1)[:n]
        return text, 5
    if kind == "period1000":
        g = torch.Generator(device="cpu").manual_seed(1)
        blk = torch.randint(1, 5, (1000,), generator=g, dtype=torch.uint8).to(dev)
        text[:n] = blk.repeat(n // 1000 + 1)[:n]
        return text, 5
    if kind == "fib":
        a, b = np.array([1], np.uint8), np.array([1, 2], np.uint8)
        while len(b) < n:
            a, b = b, np.concatenate([b, a])
        text[:n] = torch.from_numpy(b[:n].copy()).to(dev)
        return text, 3
    raise ValueError(kind)


def device_view(ptr, count, itemsize, device=0):
    """Zero-copy torch view of a device array owned by an index (uint8 / int32)."""
    import torch
    iface = {"shape": (count,), "typestr": {1: "|u1", 4: "<i4"}[itemsize], "data": (ptr, False), "version": 2}

    class Holder:
        __cuda_array_interface__ = iface
    return torch.as_tensor(Holder(), device=torch.device("cuda", device))


def check_suffix_array(text, sa, n, chunk=1 << 27):
    """The linear-time suffix-array checker on the GPU, independent of the product kernels: SA is a
    permutation of 0..n, and for every r >= 1 suffix SA[r-1] < suffix SA[r] (first symbols compared,
    ties resolved through the inverse permutation at the NEXT positions).  Together the two facts pin
    SA uniquely: it is the array the reference's constructors produce (stralg/suffix_array.c:26-48)."""
    import torch
    dev = text.device
    length = n + 1
    isa = torch.empty(length, dtype=torch.int32, device=dev)
    for lo in range(0, length, chunk):
        hi = min(length, lo + chunk)
        s = sa[lo:hi].long() & 0xFFFFFFFF
        if int(s.max()) > n:
            return False, "entry out of range"
        isa[s] = torch.arange(lo, hi, device=dev, dtype=torch.int64).to(torch.int32)
    for lo in range(0, length, chunk):
        hi = min(length, lo + chunk)
        s = sa[lo:hi].long() & 0xFFFFFFFF
        if not bool(((isa[s].long() & 0xFFFFFFFF) == torch.arange(lo, hi, device=dev)).all()):
            return False, "not a permutation"
    if int(sa[0].long() & 0xFFFFFFFF) != n:
        return False, "SA[0] != n"
    for lo in range(1, length, chunk):
        hi = min(length, lo + chunk)
        a = sa[lo - 1:hi - 1].long() & 0xFFFFFFFF
        b = sa[lo:hi].long() & 0xFFFFFFFF
        ta, tb = text[a], text[b]
        ra = isa[torch.clamp(a + 1, max=n)].long() & 0xFFFFFFFF
        rb = isa[torch.clamp(b + 1, max=n)].long() & 0xFFFFFFFF
        ok = (ta < tb) | ((ta == tb) & (ra < rb))
        if not bool(ok.all()):
            return False, f"suffixes out of order near row {lo + int((~ok).nonzero()[0])}"
        del a, b, ta, tb, ra, rb, ok
    return True, "ok"
