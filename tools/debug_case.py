"""Debug helper (run under gpurun): build one of the non-uniform test texts in a given mode and print where the
suffix array differs from the reference's."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
name, mode = sys.argv[1], sys.argv[2]
import test_gpu_nonuniform as T  # noqa: E402
import _oracle  # noqa: E402
for k, v in T.MODES[mode].items():
    os.environ[k] = v
import stralg_b200  # noqa: E402
_, sym, sigma = T._texts()[name]
codes = np.concatenate([np.asarray(sym, dtype=np.uint8), np.zeros(1, np.uint8)])
idx = stralg_b200.SuffixArrayIndex.build(codes[:-1], sigma)
sa = idx.sa()
ref = _oracle.Ref() if _oracle.Ref.available() else None
exp = ref.sa(codes, sigma, "sa_is")
bad = np.nonzero(sa != exp)[0]
print("stats", idx.stats())
print("n", len(codes) - 1, "mismatching rows", len(bad), bad[:40])
for r in bad[:12]:
    a, b = int(sa[r]), int(exp[r])
    print(r, "got", a, codes[a:a + 40].tolist(), "exp", b, codes[b:b + 40].tolist())
