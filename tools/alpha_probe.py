"""Build time of uniform random texts over alphabets that do not fill their symbol width (run under gpurun):
    python tools/alpha_probe.py N nsym[,nsym...]"""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stralg_b200
from stralg_b200 import texts as T
lib = stralg_b200.load()
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 26
for nsym in [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "3,4,5,12,16,17,20,100,255").split(",")]:
    text = T.random_codes(lib, n, nsym, T.SEED)
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        idx = stralg_b200.SuffixArrayIndex.build(text[:n], nsym + 1, bwt=True, occ=True, profile=True)
        torch.cuda.synchronize(); dt = (time.time() - t0) * 1e3
        st = idx.stats(); agg = {}
        for nm, ms, by in idx.profile():
            a = agg.setdefault(nm, [0, 0.0]); a[0] += 1; a[1] += ms
        if rep:
            print(json.dumps({"nsym": nsym, "n": n, "wall_ms": round(dt, 1), "Mchar_s": round(n / dt / 1e3, 1),
                  "stats": {k: st[k] for k in ("rounds", "k0", "round0_mode", "bucket_bits", "dense_keys", "shallow_buckets", "sorted_total")},
                  "stages": {k: [v[0], round(v[1], 2)] for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:6]}}), flush=True)
        if rep and len(sys.argv) > 3 and sys.argv[3] == "check":
            lib.b200sa_release_workspace(0)
            sa = T.device_view(idx.device_ptr("sa"), n + 1, 4)
            print("  check:", T.check_suffix_array(text, sa, n)[1], flush=True)
        idx.close()
    del text
