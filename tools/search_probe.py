"""Search-kernel throughput probe (development aid, run under gpurun)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stralg_b200  # noqa: E402

lib = stralg_b200.load()
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 28
nreads = int(float(sys.argv[2])) if len(sys.argv) > 2 else 20_000_000
m = 100
text = torch.empty(n + 1, dtype=torch.uint8, device="cuda")
lib.b200sa_synth_codes(C.c_void_p(text.data_ptr()), n, 4, 12345, 0, None)
tc = os.environ.get('PROBE_TEXTCMP', '1') == '1'
kt = os.environ.get('PROBE_KTABLE', '1') == '1'
idx = stralg_b200.SuffixArrayIndex.build(text[:n], 5, occ=True, drop_sa=not tc, textcmp=tc, ktable=kt)
reads = torch.empty(nreads * m, dtype=torch.uint8, device="cuda")
lib.b200sa_synth_reads(C.c_void_p(text.data_ptr()), n, 4, C.c_void_p(reads.data_ptr()), nreads, m, 102, 7, 0, None)
L = torch.empty(nreads, dtype=torch.int32, device="cuda")
R = torch.empty(nreads, dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    idx.search_device(reads, None, m, nreads, L, R, st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    idx.search_device(reads, None, m, nreads, L, R, st)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"textcmp={tc} generic={os.environ.get('B200SA_SEARCH_GENERIC')}: n={n} reads={nreads} {ms:.2f} ms -> {nreads/ms/1e3:.1f} M reads/s "
      f"hits={(R>L).float().mean().item():.3f} chk={int(L.long().sum())}")
