"""Per-round stage times of one build (run under gpurun): python tools/round_trace.py N kind"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import stralg_b200
from nonuniform_probe import make
n = int(float(sys.argv[1])); kind = sys.argv[2]
text, sigma, info = make(kind, n)
for rep in range(2):
    idx = stralg_b200.SuffixArrayIndex.build(text[:n], sigma, occ=True, profile=True)
    torch.cuda.synchronize()
    if rep:
        line = []
        for name, ms, by in idx.profile():
            if name == "round_keys":
                print("  ".join(line)); line = []
            line.append(f"{name}={ms:.2f}({by/1e6:.0f}MB)")
        print("  ".join(line))
        print(idx.stats())
    idx.close()
