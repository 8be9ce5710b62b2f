"""One pass over the SURVEY 8f kernels (sampled-SA build + locate, approximate search) on a
synthetic DNA index, for ncu captures:  ncu --set full -k regex:'ssa_|approx_' python tools/extras_probe.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import stralg_b200

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 28
nreads = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
napprox = int(sys.argv[3]) if len(sys.argv) > 3 else 50_000
m = 100
lib = stralg_b200.load()
text = torch.empty(n + 1, dtype=torch.uint8, device="cuda")
lib.b200sa_synth_codes(C.c_void_p(text.data_ptr()), n, 4, 1, 0, None)
reads = torch.empty(nreads * m, dtype=torch.uint8, device="cuda")
lib.b200sa_synth_reads(C.c_void_p(text.data_ptr()), n, 4, C.c_void_p(reads.data_ptr()), nreads, m, 102, 2, 0, None)
idx = stralg_b200.SuffixArrayIndex.build(text[:n], 5, ktable=True)
L = torch.empty(nreads, dtype=torch.int32, device="cuda")
R = torch.empty(nreads, dtype=torch.int32, device="cuda")
idx.search_device(reads, None, m, nreads, L, R)
idx.sample_sa(32, drop_sa=True)
poff = torch.empty(nreads + 1, dtype=torch.int64, device="cuda")
total = idx.locate_device(L, R, nreads, poff)
pos = torch.empty(max(total, 1), dtype=torch.int32, device="cuda")
idx.locate_device(L, R, nreads, poff, pos, total)
torch.cuda.synchronize()
print("located", total, "positions through the sampled SA")
rev = stralg_b200.SuffixArrayIndex.build(torch.flip(text[:n], dims=[0]).contiguous(), 5, drop_sa=True)
h = reads[: napprox * m].cpu().numpy()
r = idx.approx_search(h, fixed_len=m, max_edits=1, rev=rev)
print("approx d=1:", napprox, "reads,", len(r["L"]), "intervals")
