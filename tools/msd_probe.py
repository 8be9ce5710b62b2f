"""Development aid (run under gpurun): build the same text with the MSD and the LSD round 0,
compare the device-resident SA / BWT bit for bit and print the per-stage device times."""
import ctypes as C
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stralg_b200  # noqa: E402

lib = stralg_b200.load()


def view(ptr, n, dtype):
    """wrap a device pointer without copying"""
    cai = {"shape": (n,), "typestr": {torch.int32: "<i4", torch.uint8: "|u1"}[dtype], "data": (ptr, False), "version": 2}

    class Holder:
        __cuda_array_interface__ = cai
    return torch.as_tensor(Holder(), device="cuda")


def build(text, n, nsym, mode, reps=2):
    os.environ.pop("B200SA_ROUND0", None)
    if mode == "lsd":
        os.environ["B200SA_ROUND0"] = "lsd"
    for rep in range(reps):
        torch.cuda.synchronize()
        t0 = time.time()
        idx = stralg_b200.SuffixArrayIndex.build(text[:n], nsym + 1, profile=True, bwt=True, occ=True)
        torch.cuda.synchronize()
        dt = time.time() - t0
        if rep < reps - 1:
            idx.close()
    prof = idx.profile()
    st = idx.stats()
    agg = {}
    for name, ms, by in prof:
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += ms
        a[2] += by
    tot = sum(a[1] for a in agg.values())
    print(f"[{mode}] n={n} nsym={nsym} wall={dt*1e3:.1f} ms stages={tot:.1f} ms -> {n/dt/1e6:.0f} Mchar/s stats={st}")
    for name, (cnt, ms, by) in agg.items():
        print(f"   {name:16s} x{cnt:<3d} {ms:9.3f} ms  {by/ms/1e6 if ms else 0:8.1f} GB/s (algorithmic)")
    sys.stdout.flush()
    return idx


def run(n, nsym):
    text = torch.empty(n + 1, dtype=torch.uint8, device="cuda")
    assert lib.b200sa_synth_codes(C.c_void_p(text.data_ptr()), n, nsym, 12345, 0, None) == 0
    torch.cuda.synchronize()
    a = build(text, n, nsym, "msd")
    sa_a = view(a.device_ptr("sa"), n + 1, torch.int32).clone()
    bwt_a = view(a.device_ptr("bwt"), n + 1, torch.uint8).clone()
    pa = a.primary
    a.close()
    if n <= (1 << 31) and not MSD_ONLY:
        b = build(text, n, nsym, "lsd")
        sa_b = view(b.device_ptr("sa"), n + 1, torch.int32)
        bwt_b = view(b.device_ptr("bwt"), n + 1, torch.uint8)
        ok_sa = bool(torch.equal(sa_a, sa_b))
        ok_bwt = bool(torch.equal(bwt_a, bwt_b))
        print(f"   SA equal: {ok_sa}  BWT equal: {ok_bwt}  primary {pa} vs {b.primary}")
        if not ok_sa:
            bad = torch.nonzero(sa_a != sa_b)[:8].flatten().tolist()
            print("   first SA mismatches at", bad, [int(sa_a[i]) for i in bad], [int(sa_b[i]) for i in bad])
        b.close()
    sys.stdout.flush()


MSD_ONLY = len(sys.argv) > 3 and sys.argv[3] == "msd"

if __name__ == "__main__":
    sizes = [int(float(x)) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1 << 24]
    nsym = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    for n in sizes:
        run(n, nsym)
