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
