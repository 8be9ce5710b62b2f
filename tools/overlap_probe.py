"""Where does the end-to-end overlap go?  D2H of a 12 GB suffix array alone, next to an H2D copy,
next to a device-resident build, next to a host-text build (run under gpurun)."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import stralg_b200

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 3_000_000_000
lib = stralg_b200.load()
dev = torch.device("cuda", 0)
text = torch.empty(n + 1, dtype=torch.uint8, device=dev)
lib.b200sa_synth_codes(C.c_void_p(text.data_ptr()), n, 4, 1, 0, None)
h_text = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h_text.copy_(text[:n])
h_sa = torch.empty(n + 1, dtype=torch.int32, pin_memory=True)
idx = stralg_b200.SuffixArrayIndex.build(text[:n], 5)
cs = torch.cuda.Stream(device=dev)
hs = torch.cuda.Stream(device=dev)
d_tmp = torch.empty(n, dtype=torch.uint8, device=dev)


def d2h():
    stralg_b200._lib.check(lib.b200sa_copy_async(idx._h, 0, C.c_void_p(h_sa.data_ptr()), C.c_void_p(cs.cuda_stream)))


def timed(label, other):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(cs)
    d2h()
    e1.record(cs)
    r = other() if other else None
    t_other = time.perf_counter() - t0
    torch.cuda.synchronize()
    print(f"{label}: D2H {e0.elapsed_time(e1):.1f} ms; other work returned after {t_other * 1e3:.1f} ms; all done after "
          f"{(time.perf_counter() - t0) * 1e3:.1f} ms", flush=True)
    if r is not None:
        r.close()


def h2d():
    with torch.cuda.stream(hs):
        d_tmp.copy_(h_text, non_blocking=True)
    return None


for rep in range(2):
    timed("alone", None)
    timed("with H2D of 3 GB on another stream", h2d)
    timed("with a device-text build", lambda: stralg_b200.SuffixArrayIndex.build(text[:n], 5))
    timed("with a host-text build", lambda: stralg_b200.SuffixArrayIndex.build(h_text.numpy(), 5))
