"""CPU tests of the drop-in boundary: the shared library loads and exports every symbol that
include/b200sa.h declares, and fails loudly (no fallback) without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200sa.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200sa_[A-Za-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import stralg_b200
    from stralg_b200 import _lib
    lib = stralg_b200.load()
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/b200sa.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names


def test_no_cpu_fallback_without_device():
    import stralg_b200
    lib = stralg_b200.load()
    if lib.b200sa_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(stralg_b200.B200saError) as e:
        stralg_b200.SuffixArrayIndex.build(np.array([1, 2, 1], dtype=np.uint8), 5)
    assert e.value.code == 1  # B200SA_ERR_CUDA


def test_argument_validation():
    import stralg_b200
    lib = stralg_b200.load()
    err = C.c_int(0)
    buf = np.array([1], dtype=np.uint8)
    assert not lib.b200sa_build(C.c_void_p(buf.ctypes.data), 1, 0, 0, 0, None, C.byref(err))
    assert err.value == 2
    assert not lib.b200sa_build(C.c_void_p(buf.ctypes.data), 2 ** 32 - 1, 5, 0, 0, None, C.byref(err))
    assert err.value == 4
    assert b"2^32" in lib.b200sa_last_error()


def test_remap_table_matches_oracle(oracle):
    from stralg_b200 import RemapTable
    for raw in (b"mississippi", b"acagtgtaac", b"the quick brown fox", bytes(range(1, 120))):
        t = RemapTable(raw)
        codes, sigma, table = oracle.remap(raw)
        assert t.alphabet_size == sigma
        assert np.array_equal(t.table, table)
        assert np.array_equal(t.remap(raw), codes[:-1])
        assert t.rev_remap(t.remap(raw)) == raw
    assert RemapTable(b"acgt").remap(b"acgx") is None  # remap.c:80-84


def test_stats_struct_layout_matches_the_header(tmp_path):
    """struct b200sa_stats as the C compiler lays it out (include/b200sa.h) against the ctypes mirror
    (stralg_b200/_lib.py): same size, same field offsets -- the struct grows with the library."""
    import subprocess
    from stralg_b200._lib import Stats
    fields = [name for name, _ in Stats._fields_]
    src = tmp_path / "layout.c"
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{os.path.join(ROOT, "include", "b200sa.h")}"',
             'int main(void) {', '  printf("%zu\\n", sizeof(struct b200sa_stats));']
    lines += [f'  printf("%zu\\n", offsetof(struct b200sa_stats, {f}));' for f in fields]
    lines += ['  return 0;', '}']
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c11", "-o", str(exe), str(src)], check=True)
    out = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert out[0] == C.sizeof(Stats)
    assert out[1:] == [getattr(Stats, f).offset for f in fields]
