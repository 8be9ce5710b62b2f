"""Build-time probe on non-uniform texts (run under gpurun):
    python tools/nonuniform_probe.py N kind[,kind...] [check]
kinds: random, repeat (SURVEY 8(d) C3 repeat-rich), hg38tile (genome sample tiled with 1.6 % mutations),
       hg38tile256 (0.4 % mutations), dnan (ACGT + 5 % N in runs), uniformNN (NN equally likely letters),
       byte, unary, acgt4, period1000, fib
Prints one JSON line per text: build ms, rounds, stage times; `check` verifies the suffix array with the
checker of stralg_b200/texts.py."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stralg_b200  # noqa: E402
from stralg_b200 import texts as T  # noqa: E402

lib = stralg_b200.load()


def make(kind, n):
    if kind == "random":
        return T.random_codes(lib, n, 4, T.SEED), 5, {}
    if kind == "repeat":
        t = T.random_codes(lib, n, 4, T.SEED)
        info = T.add_repeats(t, n)
        return t, 5, info
    if kind == "hg38tile":
        return T.hg38_like(n, mut_inv=64), 5, {"mut_inv": 64}
    if kind == "hg38tile256":
        return T.hg38_like(n, mut_inv=256), 5, {"mut_inv": 256}
    if kind == "dnan":  # A C G N T: 5 % N in long runs (assembly gaps) plus a sprinkle of single N
        t, frac = T.dna_with_n(lib, n)
        return t, 6, {"N_fraction": round(frac, 4)}
    if kind.startswith("uniform"):  # uniformNN: NN equiprobable letters
        k = int(kind[7:])
        return T.random_codes(lib, n, k, T.SEED), k + 1, {}
    t, sigma = T.stress_text(lib, kind, n)
    return t, sigma, {}


if __name__ == "__main__":
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 24
    kinds = sys.argv[2].split(",") if len(sys.argv) > 2 else ["random", "repeat", "hg38tile"]
    check = len(sys.argv) > 3 and sys.argv[3] == "check"
    reps = int(os.environ.get("PROBE_REPS", "2"))
    for kind in kinds:
        text, sigma, info = make(kind, n)
        torch.cuda.synchronize()
        best = None
        for rep in range(reps):
            t0 = time.time()
            idx = stralg_b200.SuffixArrayIndex.build(text[:n], sigma, occ=True, profile=True)
            torch.cuda.synchronize()
            dt = time.time() - t0
            st = idx.stats()
            agg = {}
            for name, ms, by in idx.profile():
                a = agg.setdefault(name, [0, 0.0])
                a[0] += 1
                a[1] += ms
            rec = {"kind": kind, "n": n, "sigma": sigma, "wall_ms": round(dt * 1e3, 1),
                   "Mchar_s": round(n / dt / 1e6, 1), "rounds": st["rounds"], "k0": st["k0"],
                   "round0_mode": st["round0_mode"], "passes0": st["passes0"], "bucket_bits": st["bucket_bits"],
                   "shallow_buckets": st["shallow_buckets"], "shallow_elems": st["shallow_elems"],
                   "chain_rounds": st["chain_rounds"], "chain_elems": st["chain_elems"], "lazy": st["lazy_lookups"],
                   "resolved_small": st["resolved_small"], "small_path_elems": st["small_path_elems"],
                   "pair_placed": st["pair_placed"], "dense_keys": st["dense_keys"],
                   "pivot_rounds": st["pivot_rounds"], "pivot_elems_x": round(st["pivot_elems"] / (n + 1), 2),
                   "sorted_total_x": round(st["sorted_total"] / (n + 1), 2),
                   "stages_ms": {k: [v[0], round(v[1], 2)] for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])},
                   "stages_total_ms": round(sum(v[1] for v in agg.values()), 1), "info": info}
            if best is None or dt * 1e3 < best["wall_ms"]:
                best = rec
            if rep + 1 < reps:
                idx.close()
        if check:
            lib.b200sa_release_workspace(0)
            sa = T.device_view(idx.device_ptr("sa"), n + 1, 4)
            best["sa_ok"] = T.check_suffix_array(text, sa, n)[1]
        idx.close()
        print(json.dumps(best), flush=True)
        del text
        torch.cuda.empty_cache()
