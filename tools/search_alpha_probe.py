"""Exact-search throughput on alphabets other than ACGT (generic O layout; run under gpurun):
    python tools/search_alpha_probe.py N NREADS kind   (kind: dnan | uniformNN)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import stralg_b200
from nonuniform_probe import make
n = int(float(sys.argv[1])); nreads = int(float(sys.argv[2])); kind = sys.argv[3]; m = 100
text, sigma, info = make(kind, n)
tcmp = os.environ.get('PROBE_TEXTCMP', '1') == '1'
ktab = os.environ.get('PROBE_KTABLE', '1') == '1'
idx = stralg_b200.SuffixArrayIndex.build(text[:n], sigma, occ=True, textcmp=tcmp, ktable=ktab)
g = torch.Generator(device="cuda"); g.manual_seed(1)
starts = torch.randint(0, n - m, (nreads,), generator=g, device="cuda")
reads = text[(starts[:, None] + torch.arange(m, device="cuda")[None, :])].contiguous()
if kind == "dnan" and os.environ.get("PROBE_READS_WITH_N", "0") != "1":
    # a sequencer does not read assembly gaps: reads that hold an N are drawn again (PROBE_READS_WITH_N=1 keeps them;
    # an all-N read never narrows its interval and holds its whole warp for 100 steps)
    for _ in range(8):
        bad = (reads == 4).any(dim=1)
        nb = int(bad.sum())
        if not nb:
            break
        st2 = torch.randint(0, n - m, (nb,), generator=g, device="cuda")
        reads[bad] = text[(st2[:, None] + torch.arange(m, device="cuda")[None, :])]
miss = torch.arange(nreads, device="cuda") % 10 == 0
reads[miss] = torch.randint(1, sigma, (int(miss.sum()), m), generator=g, device="cuda", dtype=torch.int16).to(torch.uint8)
L = torch.empty(nreads, dtype=torch.int32, device="cuda"); R = torch.empty_like(L)
for _ in range(2):
    idx.search_device(reads.view(-1), None, m, nreads, L, R)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    idx.search_device(reads.view(-1), None, m, nreads, L, R)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
found = int(((R.long() & 0xffffffff) > (L.long() & 0xffffffff)).sum())
print({"kind": kind, "n": n, "sigma": sigma, "reads": nreads, "ms": round(ms, 2), "Mreads_s": round(nreads / ms / 1e3, 1), "found": found, "occ_layout": idx.stats()["occ_layout"], "textcmp": tcmp, "ktable_k": idx.stats()["ktable_k"]})
