import ctypes as C, sys, os
sys.path.insert(0, os.getcwd())
import torch, stralg_b200
lib = stralg_b200.load()
n = 3_000_000_000
text = torch.empty(n + 1, dtype=torch.uint8, device="cuda")
lib.b200sa_synth_codes(C.c_void_p(text.data_ptr()), n, 4, 12345, 0, None)
idx = stralg_b200.SuffixArrayIndex.build(text[:n], 5, occ=True, textcmp=True, ktable=True, profile=True)
print({k: round(v, 2) for k, v, _ in idx.profile() if k in ("ktable", "inverse", "occ_build")})
