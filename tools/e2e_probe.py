"""Development aid (run under gpurun): where the host-buffer (e2e) build spends its time --
pinned H2D of the text, device build, D2H of SA / O -- and what the PCIe link gives for plain copies."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stralg_b200  # noqa: E402

lib = stralg_b200.load()
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 3_000_000_000


def t(fn, reps=2):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


text = torch.empty(n + 1, dtype=torch.uint8, device="cuda")
assert lib.b200sa_synth_codes(C.c_void_p(text.data_ptr()), n, 4, 1, 0, None) == 0
h_text = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h_text.copy_(text[:n])
h_sa = torch.empty(n + 1, dtype=torch.int32, pin_memory=True)
d_sa = torch.empty(n + 1, dtype=torch.int32, device="cuda")
dt = t(lambda: text[:n].copy_(h_text, non_blocking=True))
print(f"plain H2D {n/1e9:.1f} GB: {dt*1e3:.1f} ms = {n/dt/1e9:.1f} GB/s")
dt = t(lambda: h_sa.copy_(d_sa, non_blocking=True))
print(f"plain D2H {4*n/1e9:.1f} GB: {dt*1e3:.1f} ms = {4*n/dt/1e9:.1f} GB/s")
del d_sa
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    idx = stralg_b200.SuffixArrayIndex.build(h_text.numpy(), 5, occ=True)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    stralg_b200._lib.check(lib.b200sa_copy_sa(idx._h, C.c_void_p(h_sa.data_ptr())))
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    occ_bytes = idx.stats()["occ_bytes"]
    h_occ = torch.empty(occ_bytes, dtype=torch.uint8, pin_memory=True) if rep == 0 else h_occ
    t2b = time.perf_counter()
    stralg_b200._lib.check(lib.b200sa_copy_occ(idx._h, C.c_void_p(h_occ.data_ptr())))
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    idx.close()
    print(f"rep {rep}: build-from-host {1e3*(t1-t0):.1f} ms, copy_sa {1e3*(t2-t1):.1f} ms, copy_occ {1e3*(t3-t2b):.1f} ms")
