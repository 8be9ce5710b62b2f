"""Development aid (run under gpurun): search throughput on the 3 Gbp index with and without the
k-mer seed table / text-compare shortcut, results compared bit for bit."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stralg_b200  # noqa: E402

lib = stralg_b200.load()
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 3_000_000_000
reads_n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 100_000_000
m = 100
text = torch.empty(n + 1, dtype=torch.uint8, device="cuda")
assert lib.b200sa_synth_codes(C.c_void_p(text.data_ptr()), n, 4, 88172645463325252, 0, None) == 0
idx = stralg_b200.SuffixArrayIndex.build(text[:n], 5, occ=True, textcmp=True, ktable=True)
lib.b200sa_release_workspace(0)
reads = torch.empty(reads_n * m, dtype=torch.uint8, device="cuda")
assert lib.b200sa_synth_reads(C.c_void_p(text.data_ptr()), n, 4, C.c_void_p(reads.data_ptr()), reads_n, m, 102, 7, 0, None) == 0
ref = None
for name, env in [("ktable+textcmp", {}), ("textcmp only", {"B200SA_SEARCH_NO_KTABLE": "1"})]:
    pass
L = torch.empty(reads_n, dtype=torch.int32, device="cuda")
R = torch.empty(reads_n, dtype=torch.int32, device="cuda")


def run(label):
    for _ in range(2):
        idx.search_device(reads, None, m, reads_n, L, R, 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        idx.search_device(reads, None, m, reads_n, L, R, 0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    counts = (C.c_uint64 * 4)()
    lib.b200sa_search_traffic(idx._h, C.c_void_p(reads.data_ptr()), None, m, reads_n, C.c_void_p(L.data_ptr()),
                              C.c_void_p(R.data_ptr()), counts, None)
    print(f"{label}: {ms:.2f} ms -> {reads_n/ms/1e6:.3f} G reads/s; per read: blocks {counts[0]/reads_n:.2f} "
          f"pattern words {counts[1]/reads_n:.2f} text words {counts[2]/reads_n:.2f} 4-byte loads {counts[3]/reads_n:.2f}")
    return L.clone(), R.clone()


a = run("ktable + textcmp")
print("stats", idx.stats())
