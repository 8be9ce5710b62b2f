"""Config-5 stress: SA (+LCP) construction on byte-alphabet and periodic texts (run under gpurun).
Checks the suffix array with the permutation + adjacent-order property (torch, chunked)."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stralg_b200  # noqa: E402

lib = stralg_b200.load()
dev = torch.device("cuda", 0)
CHUNK = 1 << 27


def view(ptr, count, itemsize):
    iface = {"shape": (count,), "typestr": {1: "|u1", 4: "<i4"}[itemsize], "data": (ptr, False), "version": 2}

    class Holder:
        __cuda_array_interface__ = iface
    return torch.as_tensor(Holder(), device=dev)


def check_sa(text, sa, n):
    length = n + 1
    isa = torch.empty(length, dtype=torch.int32, device=dev)
    for lo in range(0, length, CHUNK):
        hi = min(length, lo + CHUNK)
        s = sa[lo:hi].long() & 0xFFFFFFFF
        isa[s] = torch.arange(lo, hi, device=dev, dtype=torch.int64).to(torch.int32)
    for lo in range(0, length, CHUNK):
        hi = min(length, lo + CHUNK)
        s = sa[lo:hi].long() & 0xFFFFFFFF
        assert bool(((isa[s].long() & 0xFFFFFFFF) == torch.arange(lo, hi, device=dev)).all()), "not a permutation"
    for lo in range(1, length, CHUNK):
        hi = min(length, lo + CHUNK)
        a = sa[lo - 1:hi - 1].long() & 0xFFFFFFFF
        b = sa[lo:hi].long() & 0xFFFFFFFF
        ta, tb = text[a], text[b]
        ra = isa[torch.clamp(a + 1, max=n)].long() & 0xFFFFFFFF
        rb = isa[torch.clamp(b + 1, max=n)].long() & 0xFFFFFFFF
        ok = (ta < tb) | ((ta == tb) & (ra < rb))
        assert bool(ok.all()), "suffixes out of order"
    return True


def make(kind, n):
    text = torch.zeros(n + 1, dtype=torch.uint8, device=dev)
    if kind == "byte":
        lib.b200sa_synth_codes(C.c_void_p(text.data_ptr()), n, 255, 5, 0, None)
        return text, 256
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


if __name__ == "__main__":
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 24
    kinds = sys.argv[2].split(",") if len(sys.argv) > 2 else ["byte", "unary", "acgt4", "period1000", "fib"]
    lcp = len(sys.argv) > 3 and sys.argv[3] == "lcp"
    for kind in kinds:
        text, sigma = make(kind, n)
        torch.cuda.synchronize()
        t0 = time.time()
        idx = stralg_b200.SuffixArrayIndex.build(text[:n], sigma, occ=False, lcp=lcp, profile=True)
        torch.cuda.synchronize()
        dt = time.time() - t0
        st = idx.stats()
        agg = {}
        for name, ms, by in idx.profile():
            agg[name] = agg.get(name, 0.0) + ms
        top = sorted(agg.items(), key=lambda kv: -kv[1])[:9]
        top.append(('ALL_STAGES', sum(agg.values())))
        lib.b200sa_release_workspace(0)
        sa = view(idx.device_ptr("sa"), n + 1, 4)
        ok = check_sa(text, sa, n)
        print(f"{kind:11s} n={n} sigma={sigma} lcp={lcp} build={dt*1e3:9.1f} ms ({n/dt/1e6:8.1f} Mchar/s) rounds={st['rounds']} "
              f"k0={st['k0']} sorted_total={st['sorted_total']/ (n+1):.1f}x ok={ok} top={[(k, round(v,1)) for k,v in top]}",
              flush=True)
        idx.close()
        del text, sa
        torch.cuda.empty_cache()
