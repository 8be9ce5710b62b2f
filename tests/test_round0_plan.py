"""Host-side checks of the round-0 plan (stralg_b200/csrc/round0_msd.cu: msd_make_plan) through the diagnostic
entry point b200sa_plan_round0 -- no GPU needed.  The plan decides how a suffix's first K symbols become a key, how
the key is cut into bucket digits and what is left for the 8-byte element; a plan that breaks one of its own limits
shows only at particular sizes (3 Gbp cannot be rebuilt in every test run), so the limits are swept here."""
import ctypes as C
import math

import numpy as np
import pytest

import stralg_b200

FIELDS = ("ok", "levels", "D1", "D2", "D3", "BB", "K", "KB", "pb", "R", "dense", "Khi", "Klo", "powlo", "bits", "_")


def plan(lib, length, sigma, counts=None):
    out = (C.c_uint64 * 16)()
    cp = None
    if counts is not None:
        arr = np.zeros(256, dtype=np.uint64)
        arr[:len(counts)] = counts
        cp = arr.ctypes.data_as(C.c_void_p)
        plan.keep = arr
    assert lib.b200sa_plan_round0(length, sigma, cp, out) == 0
    return dict(zip(FIELDS, [int(v) for v in out]))


@pytest.fixture(scope="module")
def lib():
    return stralg_b200.load()


LENGTHS = [2, 3, 17, 1000, 65537, 1 << 20, (1 << 24) + 1, 1 << 28, (1 << 30) + 1, 3_000_000_001, (1 << 32) - 1]
SIGMAS = [2, 3, 4, 5, 6, 7, 12, 15, 16, 17, 18, 21, 33, 101, 129, 201, 231, 232, 256]


@pytest.mark.parametrize("sigma", SIGMAS)
def test_plan_respects_its_limits(lib, sigma):
    nsym = sigma - 1
    for length in LENGTHS:
        p = plan(lib, length, sigma)
        b = p["bits"]
        assert b == (1 if nsym <= 2 else 2 if nsym <= 4 else 4 if nsym <= 16 else 8)
        if not p["ok"]:
            continue
        D = [p["D1"], p["D2"], p["D3"]]
        assert 1 <= p["levels"] <= 3 and all(0 < d <= 10 for d in D[:p["levels"]]) and all(d == 0 for d in D[p["levels"]:])
        assert sum(D) == p["BB"] and p["R"] == p["KB"] - p["BB"] and 0 <= p["R"] <= 32
        # the element: 32 bits of suffix start, the preceding symbol, the key without its first digit
        assert p["pb"] in (0, b) and p["KB"] - p["D1"] <= 32 - p["pb"], (length, p)
        if p["dense"]:
            assert p["dense"] == nsym and p["Khi"] + p["Klo"] == p["K"]
            assert p["K"] * b + p["pb"] <= 128                       # the window of the text a key is read from
            span = nsym ** p["K"]
            assert p["KB"] == (span - 1).bit_length() and p["KB"] <= 42
            assert p["powlo"] == nsym ** p["Klo"] < 1 << 32 and nsym ** p["Khi"] < 1 << 32
            # enough used buckets for the in-SM sort: an average bucket of equally likely letters fits
            used = ((span - 1) >> p["R"]) + 1
            assert length / used <= 3328 or p["BB"] == 30 or p["BB"] == p["KB"], (length, p)
        else:
            assert p["KB"] == p["K"] * b and p["KB"] + p["pb"] <= 64 and all(d % b == 0 for d in D)


def test_dense_keys_only_for_sparse_alphabets(lib):
    for sigma, dense in ((3, False), (4, True), (5, False), (6, True), (15, True), (16, False), (17, False), (18, True),
                         (21, True), (201, True), (232, False), (256, False)):
        p = plan(lib, 1 << 26, sigma)
        assert p["ok"] and bool(p["dense"]) == dense, (sigma, p)


def test_key_depth_follows_the_letter_counts(lib):
    """DNA + N: the rare letter must not shorten the key (order-0 entropy instead of log2 of the letter count)."""
    n = 3_000_000_000
    even = plan(lib, n + 1, 6, [1] + [n // 5] * 5)
    rare_n = plan(lib, n + 1, 6, [1, n // 4, n // 4, n // 4, n // 1000, n // 4])
    assert even["dense"] == rare_n["dense"] == 5
    assert rare_n["K"] >= even["K"] and rare_n["K"] >= 18, (even, rare_n)
    assert math.log2(5) * rare_n["K"] <= 42
