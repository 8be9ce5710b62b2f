"""Stage profile of the texts the mid-size sweep found slow per symbol (run under gpurun)."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stralg_b200
dev = torch.device("cuda:0")
g = torch.Generator(device="cuda"); g.manual_seed(5)
cpu = torch.Generator(); cpu.manual_seed(5)
ri = lambda lo, hi: int(torch.randint(lo, hi, (1,), generator=cpu))
def rand_sym(n, nsym): return torch.randint(1, nsym + 1, (n,), generator=g, device=dev, dtype=torch.int16).to(torch.uint8)
def copies(n, nsym, lmax, k=300):
    t = rand_sym(n, nsym)
    for _ in range(k):
        ln = ri(1, lmax); a, b = ri(0, n - ln), ri(0, n - ln)
        t[b:b + ln] = t[a:a + ln].clone()
    return t
def fib(n):
    x, y = torch.tensor([1], dtype=torch.uint8, device=dev), torch.tensor([1, 2], dtype=torch.uint8, device=dev)
    while y.numel() < n: x, y = y, torch.cat([y, x])
    return y[:n].clone()
def tm(n):
    i = torch.arange(n, device=dev); par = torch.zeros(n, dtype=torch.int64, device=dev)
    for s in range(32): par ^= (i >> s) & 1
    return (par + 1).to(torch.uint8)
cases = {"copies20": lambda: (copies(3020890, 20, 92782), 21), "copies4": lambda: (copies(1120140, 4, 61394), 5),
         "fib": lambda: (fib(1560564), 5), "tm16": lambda: (tm(1509757), 17),
         "runs100": lambda: (torch.repeat_interleave(rand_sym(500000, 100), torch.randint(1, 30, (500000,), generator=g, device=dev))[:1403699].clone(), 101)}
for name in (sys.argv[1].split(",") if len(sys.argv) > 1 else cases):
    text, sigma = cases[name]()
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        idx = stralg_b200.SuffixArrayIndex.build(text, sigma, bwt=True, occ=True, profile=True)
        torch.cuda.synchronize(); dt = (time.time() - t0) * 1e3
        st = idx.stats(); agg = {}
        for nm, ms, by in idx.profile():
            a = agg.setdefault(nm, [0, 0.0]); a[0] += 1; a[1] += ms
        if rep: print(json.dumps({"case": name, "n": text.numel(), "wall_ms": round(dt, 1), "gpu_ms": round(sum(v[1] for v in agg.values()), 1),
              "stats": {k: st[k] for k in ("rounds", "round0_mode", "pivot_rounds", "pair_placed", "chain_rounds", "lazy_lookups", "sorted_total")},
              "stages": {k: [v[0], round(v[1], 2)] for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]}}), flush=True)
        idx.close()
