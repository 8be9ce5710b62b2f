"""Per-stage device timings of the build at a few sizes (development aid, run under gpurun)."""
import ctypes as C
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stralg_b200  # noqa: E402

lib = stralg_b200.load()


def run(n, nsym, reps=2, **kw):
    text = torch.empty(n + 1, dtype=torch.uint8, device="cuda")
    assert lib.b200sa_synth_codes(C.c_void_p(text.data_ptr()), n, nsym, 12345, 0, None) == 0
    torch.cuda.synchronize()
    for rep in range(reps):
        t0 = time.time()
        idx = stralg_b200.SuffixArrayIndex.build(text[:n], nsym + 1, profile=True, **kw)
        torch.cuda.synchronize()
        dt = time.time() - t0
        prof = idx.profile()
        st = idx.stats()
        idx.close()
    agg = {}
    for name, ms, by in prof:
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += ms
        a[2] += by
    tot = sum(a[1] for a in agg.values())
    print(f"n={n} nsym={nsym} kw={kw} wall={dt*1e3:.1f} ms  stages={tot:.1f} ms  -> {n/dt/1e6:.0f} Mchar/s  stats={st}")
    for name, (cnt, ms, by) in agg.items():
        print(f"   {name:16s} x{cnt:<3d} {ms:9.3f} ms  {by/ms/1e6 if ms else 0:8.1f} GB/s (algorithmic)")
    sys.stdout.flush()


if __name__ == "__main__":
    sizes = [int(float(x)) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1 << 24]
    nsym = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    for n in sizes:
        run(n, nsym, occ=True)
