"""ctypes bindings for the parity oracle (TEST INFRASTRUCTURE ONLY).

Two checkers live under ``oracle/``:

* ``oracle/liboracle.so``  -- our CPU restatement (``oracle/stralg_oracle.c``), always present.
* ``oracle/_ref/libstralg_ref.so`` -- the unmodified reference compiled from /root/reference
  by ``oracle/Makefile`` (travels to the GPU box as a prebuilt file; may be absent).

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl reference`` legs
import this module.  Nothing in ``stralg_b200`` does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libstralg_ref.so")

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)


def _p(a, t):
    return a.ctypes.data_as(t)


def build_oracle():
    """Compile oracle/ (and oracle/_ref when the reference tree is present)."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True, stdout=subprocess.DEVNULL)


class OracleApproxOut(C.Structure):  # struct oracle_approx_out (oracle/stralg_oracle.c)
    _fields_ = [("nhits", C.c_uint64), ("cap", C.c_uint64), ("L", u32p), ("R", u32p), ("mlen", u32p),
                ("cig_off", u64p), ("cigars", C.POINTER(C.c_char)), ("cig_bytes", C.c_uint64),
                ("cig_cap", C.c_uint64)]


class OracleRemap(C.Structure):
    _fields_ = [("alphabet_size", C.c_uint32), ("table", C.c_int16 * 256), ("rev", C.c_int16 * 256)]


class Oracle:
    """Restatement oracle: numpy in, numpy out."""

    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build_oracle()
        self.lib = C.CDLL(ORACLE_SO)
        self.lib.oracle_sa_doubling.restype = C.c_int
        self.lib.oracle_remap_apply.restype = C.c_int

    # --- remap -------------------------------------------------------------------------
    def remap(self, raw: bytes):
        """Returns (codes[n+1] uint8 with trailing 0, sigma, table[256] int16)."""
        t = OracleRemap()
        buf = np.frombuffer(raw, dtype=np.uint8)
        self.lib.oracle_remap_init(C.byref(t), _p(buf, u8p) if len(buf) else None, C.c_uint64(len(buf)))
        out = np.zeros(len(buf) + 1, dtype=np.uint8)
        rc = self.lib.oracle_remap_apply(C.byref(t), _p(buf, u8p) if len(buf) else None,
                                         C.c_uint64(len(buf)), _p(out, u8p))
        assert rc == 0
        return out, int(t.alphabet_size), np.array(list(t.table), dtype=np.int16)

    def remap_pattern(self, table: np.ndarray, raw: bytes):
        """Remap a pattern with an existing table; None if a letter has no code (remap.c:80-84)."""
        codes = table[np.frombuffer(raw, dtype=np.uint8)]
        if (codes < 0).any():
            return None
        return codes.astype(np.uint8)

    # --- suffix array family -----------------------------------------------------------
    def sa(self, codes: np.ndarray, method: str = "auto") -> np.ndarray:
        """codes: uint8[n+1] with trailing 0.  Returns SA uint32[n+1]."""
        n = len(codes) - 1
        sa = np.empty(n + 1, dtype=np.uint32)
        if method == "qsort":
            self.lib.oracle_sa_qsort(_p(codes, u8p), C.c_uint32(n), _p(sa, u32p))
        else:
            rc = self.lib.oracle_sa_doubling(_p(codes, u8p), C.c_uint32(n), _p(sa, u32p))
            assert rc == 0
        return sa

    def inverse(self, sa):
        isa = np.empty_like(sa)
        self.lib.oracle_inverse(_p(sa, u32p), C.c_uint32(len(sa)), _p(isa, u32p))
        return isa

    def lcp(self, codes, sa, isa=None):
        if isa is None:
            isa = self.inverse(sa)
        lcp = np.empty_like(sa)
        self.lib.oracle_lcp_kasai(_p(codes, u8p), _p(sa, u32p), _p(isa, u32p), C.c_uint32(len(sa)),
                                  _p(lcp, u32p))
        return lcp

    # --- BWT tables --------------------------------------------------------------------
    def bwt(self, codes, sa):
        out = np.empty(len(sa), dtype=np.uint8)
        self.lib.oracle_bwt(_p(codes, u8p), _p(sa, u32p), C.c_uint32(len(sa)), _p(out, u8p))
        return out

    def c_table(self, codes, sigma):
        c = np.zeros(sigma, dtype=np.uint32)
        self.lib.oracle_c_table(_p(codes, u8p), C.c_uint32(len(codes)), C.c_uint32(sigma), _p(c, u32p))
        return c

    def o_table(self, bwt, sigma):
        """Dense O, shape (len+1, sigma): o[i, a] = #{k < i: bwt[k] == a} (bwt.c:47-65 layout)."""
        o = np.empty((len(bwt) + 1, sigma), dtype=np.uint32)
        self.lib.oracle_o_table_dense(_p(bwt, u8p), C.c_uint32(len(bwt)), C.c_uint32(sigma), _p(o, u32p))
        return o

    def o_checkpoints(self, bwt, sigma, stride=64):
        rows = len(bwt) // stride + 1
        ck = np.empty((rows, sigma), dtype=np.uint32)
        self.lib.oracle_o_checkpoints(_p(bwt, u8p), C.c_uint32(len(bwt)), C.c_uint32(sigma),
                                      C.c_uint32(stride), _p(ck, u32p))
        return ck

    def o_probe(self, bwt, ck, sigma, stride, a, i):
        a = np.ascontiguousarray(a, dtype=np.uint8)
        i = np.ascontiguousarray(i, dtype=np.uint32)
        out = np.empty(len(a), dtype=np.uint32)
        self.lib.oracle_o_probe(_p(bwt, u8p), _p(ck, u32p), C.c_uint32(sigma), C.c_uint32(stride),
                                _p(a, u8p), _p(i, u32p), C.c_uint64(len(a)), _p(out, u32p))
        return out

    # --- search ------------------------------------------------------------------------
    def search_dense(self, c, o, length, pat, off):
        npat = len(off) - 1
        L = np.empty(npat, dtype=np.uint32)
        R = np.empty(npat, dtype=np.uint32)
        pat = np.ascontiguousarray(pat, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        self.lib.oracle_search_dense(_p(c, u32p), _p(o, u32p), C.c_uint32(o.shape[1]), C.c_uint32(length),
                                     _p(pat, u8p), _p(off, u64p), C.c_uint64(npat), _p(L, u32p), _p(R, u32p))
        return L, R

    def search_ck(self, c, bwt, ck, stride, pat, off, threads=1):
        npat = len(off) - 1
        L = np.empty(npat, dtype=np.uint32)
        R = np.empty(npat, dtype=np.uint32)
        pat = np.ascontiguousarray(pat, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        self.lib.oracle_search_ck(_p(c, u32p), _p(bwt, u8p), _p(ck, u32p), C.c_uint32(ck.shape[1]),
                                  C.c_uint32(stride), C.c_uint32(len(bwt)), _p(pat, u8p), _p(off, u64p),
                                  C.c_uint64(npat), _p(L, u32p), _p(R, u32p), C.c_uint32(threads))
        return L, R

    def locate(self, sa, L, R):
        npat = len(L)
        off = np.empty(npat + 1, dtype=np.uint64)
        self.lib.oracle_locate(_p(sa, u32p), _p(L, u32p), _p(R, u32p), C.c_uint64(npat), _p(off, u64p), None)
        pos = np.empty(int(off[-1]), dtype=np.uint32)
        self.lib.oracle_locate(_p(sa, u32p), _p(L, u32p), _p(R, u32p), C.c_uint64(npat), _p(off, u64p),
                               _p(pos, u32p))
        return off, pos

    # --- approximate search (bwt.c:226-382) ---------------------------------------------
    def approx(self, c, o, ro, length, pattern, max_edits):
        """One pattern against dense tables o / ro (ro may be None: no D table).
        Returns (L, R, match_length, [cigar, ...]) in the order the reference reports intervals."""
        out = OracleApproxOut()
        pat = np.ascontiguousarray(pattern, dtype=np.uint8)
        c = np.ascontiguousarray(c, dtype=np.uint32)
        o = np.ascontiguousarray(o, dtype=np.uint32)
        rop = None
        if ro is not None:
            ro = np.ascontiguousarray(ro, dtype=np.uint32)
            rop = _p(ro, u32p)
        self.lib.oracle_approx_dense(_p(c, u32p), _p(o, u32p), rop, C.c_uint32(o.shape[1]), C.c_uint32(length),
                                     _p(pat, u8p), C.c_uint32(len(pat)), C.c_int(max_edits), C.byref(out))
        n = int(out.nhits)
        L = np.ctypeslib.as_array(out.L, shape=(n,)).copy() if n else np.zeros(0, np.uint32)
        R = np.ctypeslib.as_array(out.R, shape=(n,)).copy() if n else np.zeros(0, np.uint32)
        ml = np.ctypeslib.as_array(out.mlen, shape=(n,)).copy() if n else np.zeros(0, np.uint32)
        raw = C.string_at(out.cigars, int(out.cig_bytes)) if n else b""
        cig = [x.decode() for x in raw.split(b"\0")[:-1]]
        self.lib.oracle_approx_free(C.byref(out))
        assert len(cig) == n
        return L, R, ml, cig

    def approx_dtable(self, c, ro, length, pattern):
        pat = np.ascontiguousarray(pattern, dtype=np.uint8)
        c = np.ascontiguousarray(c, dtype=np.uint32)
        ro = np.ascontiguousarray(ro, dtype=np.uint32)
        d = np.zeros(len(pat), dtype=np.int32)
        self.lib.oracle_approx_dtable(_p(c, u32p), _p(ro, u32p), C.c_uint32(ro.shape[1]), C.c_uint32(length),
                                      _p(pat, u8p), C.c_uint32(len(pat)), d.ctypes.data_as(C.POINTER(C.c_int)))
        return d

    def random_codes(self, n, nsym=4, seed=0):
        out = np.empty(n + 1, dtype=np.uint8)
        self.lib.oracle_random_codes(_p(out, u8p), C.c_uint64(n), C.c_uint32(nsym), C.c_uint64(seed))
        return out


# ---- the real reference ----------------------------------------------------------------------
class RefSuffixArray(C.Structure):  # stralg/suffix_array.h:10-20
    _fields_ = [("string", u8p), ("length", C.c_uint32), ("array", u32p), ("inverse", u32p), ("lcp", u32p)]


class RefRemapTable(C.Structure):  # stralg/remap.h:9-19
    _fields_ = [("alphabet_size", C.c_uint32), ("table", C.c_byte * 256), ("rev_table", C.c_byte * 128)]


class RefBwtTable(C.Structure):  # stralg/bwt.h:36-44
    _fields_ = [("remap_table", C.POINTER(RefRemapTable)), ("sa", C.POINTER(RefSuffixArray)),
                ("c_table", u32p), ("o_table", u32p), ("o_indices", C.POINTER(u32p)),
                ("ro_table", u32p), ("ro_indices", C.POINTER(u32p))]


class RefExactIter(C.Structure):  # stralg/bwt.h:168-173
    _fields_ = [("sa", C.POINTER(RefSuffixArray)), ("L", C.c_uint32), ("i", C.c_int64), ("R", C.c_uint32)]


class RefExactMatch(C.Structure):  # stralg/bwt.h:180-182
    _fields_ = [("pos", C.c_uint32)]


class RefIndexVector(C.Structure):  # stralg/vectors.h:29-33
    _fields_ = [("data", u32p), ("size", C.c_uint32), ("used", C.c_uint32)]


class RefStringVector(C.Structure):  # stralg/vectors.h:153-157
    _fields_ = [("data", C.POINTER(C.c_char_p)), ("size", C.c_uint32), ("used", C.c_uint32)]


class RefApproxIter(C.Structure):  # stralg/bwt.h:246-259
    _fields_ = [("bwt_table", C.POINTER(RefBwtTable)), ("remapped_pattern", u8p),
                ("L", C.c_uint32), ("R", C.c_uint32), ("next_interval", C.c_uint32),
                ("Ls", RefIndexVector), ("Rs", RefIndexVector), ("cigars", RefStringVector),
                ("match_lengths", RefIndexVector), ("m", C.c_uint32), ("edits_buf", C.c_char_p),
                ("D_table", C.POINTER(C.c_int))]


class RefApproxMatch(C.Structure):  # stralg/bwt.h:277-281
    _fields_ = [("cigar", C.c_char_p), ("position", C.c_uint32), ("match_length", C.c_uint32)]


def bind_stralg_api(lib):
    """Declare the suffix_array.h / bwt.h / remap.h signatures on a CDLL (reference or compat)."""
    sap = C.POINTER(RefSuffixArray)
    for name in ("qsort_sa_construction", "skew_sa_construction"):
        getattr(lib, name).restype = sap
        getattr(lib, name).argtypes = [u8p]
    for name in ("sa_is_construction", "sa_is_mem_construction"):
        getattr(lib, name).restype = sap
        getattr(lib, name).argtypes = [u8p, C.c_uint32]
    lib.free_suffix_array.argtypes = [sap]
    lib.free_suffix_array.restype = None
    lib.compute_inverse.argtypes = [sap]
    lib.compute_inverse.restype = None
    lib.compute_lcp.argtypes = [sap]
    lib.compute_lcp.restype = None
    lib.alloc_remap_table.restype = C.POINTER(RefRemapTable)
    lib.alloc_remap_table.argtypes = [u8p]
    lib.free_remap_table.argtypes = [C.POINTER(RefRemapTable)]
    lib.free_remap_table.restype = None
    lib.remap.restype = C.c_void_p
    lib.remap.argtypes = [u8p, u8p, C.POINTER(RefRemapTable)]
    lib.alloc_bwt_table.restype = C.POINTER(RefBwtTable)
    lib.alloc_bwt_table.argtypes = [sap, sap, C.POINTER(RefRemapTable)]
    lib.free_bwt_table.argtypes = [C.POINTER(RefBwtTable)]
    lib.free_bwt_table.restype = None
    lib.build_complete_table.restype = C.POINTER(RefBwtTable)
    lib.build_complete_table.argtypes = [u8p, C.c_bool]
    lib.completely_free_bwt_table.argtypes = [C.POINTER(RefBwtTable)]
    lib.completely_free_bwt_table.restype = None
    lib.init_bwt_exact_match_iter.argtypes = [C.POINTER(RefExactIter), C.POINTER(RefBwtTable), u8p]
    lib.init_bwt_exact_match_iter.restype = None
    lib.next_bwt_exact_match_iter.argtypes = [C.POINTER(RefExactIter), C.POINTER(RefExactMatch)]
    lib.next_bwt_exact_match_iter.restype = C.c_bool
    if hasattr(lib, "init_bwt_approx_iter"):
        lib.init_bwt_approx_iter.argtypes = [C.POINTER(RefApproxIter), C.POINTER(RefBwtTable), u8p, C.c_int]
        lib.init_bwt_approx_iter.restype = None
        lib.next_bwt_approx_match.argtypes = [C.POINTER(RefApproxIter), C.POINTER(RefApproxMatch)]
        lib.next_bwt_approx_match.restype = C.c_bool
        lib.dealloc_bwt_approx_iter.argtypes = [C.POINTER(RefApproxIter)]
        lib.dealloc_bwt_approx_iter.restype = None
    lib.lower_bound_k.restype = C.c_uint32
    lib.lower_bound_k.argtypes = [sap, C.c_uint32, C.c_uint8, C.c_uint32, C.c_uint32]
    lib.upper_bound_k.restype = C.c_uint32
    lib.upper_bound_k.argtypes = [sap, C.c_uint32, C.c_uint8, C.c_uint32, C.c_uint32]
    lib.lower_bound_search.restype = C.c_uint32
    lib.lower_bound_search.argtypes = [sap, u8p]
    lib.upper_bound_search.restype = C.c_uint32
    lib.upper_bound_search.argtypes = [sap, u8p]
    return lib


class Ref:
    """The unmodified reference (oracle/_ref/libstralg_ref.so) behind a numpy facade."""

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    def __init__(self):
        self.lib = bind_stralg_api(C.CDLL(REF_SO))

    def sa(self, codes: np.ndarray, sigma: int, method: str = "sa_is") -> np.ndarray:
        """codes: uint8[n+1] remapped text with trailing 0."""
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        if method in ("sa_is", "sa_is_mem"):
            fn = self.lib.sa_is_construction if method == "sa_is" else self.lib.sa_is_mem_construction
            sa = fn(_p(codes, u8p), C.c_uint32(sigma))
        elif method == "skew":
            sa = self.lib.skew_sa_construction(_p(codes, u8p))
        else:
            sa = self.lib.qsort_sa_construction(_p(codes, u8p))
        out = np.ctypeslib.as_array(sa.contents.array, shape=(sa.contents.length,)).copy()
        self.lib.free_suffix_array(sa)
        return out

    def sa_lcp(self, codes, sigma):
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        sa = self.lib.sa_is_construction(_p(codes, u8p), C.c_uint32(sigma))
        self.lib.compute_lcp(sa)
        n1 = sa.contents.length
        arr = np.ctypeslib.as_array(sa.contents.array, shape=(n1,)).copy()
        isa = np.ctypeslib.as_array(sa.contents.inverse, shape=(n1,)).copy()
        lcp = np.ctypeslib.as_array(sa.contents.lcp, shape=(n1,)).copy()
        self.lib.free_suffix_array(sa)
        return arr, isa, lcp

    def tables(self, raw: bytes, include_reverse: bool = False):
        """build_complete_table (bwt.c:134-161) -> dict(codes, sigma, sa, c, o[(len+1), sigma], handle)
        (+ ro with include_reverse)."""
        buf = C.create_string_buffer(raw, len(raw) + 1)
        t = self.lib.build_complete_table(C.cast(buf, u8p), include_reverse)
        tc = t.contents
        sa = tc.sa.contents
        n1 = sa.length
        sigma = tc.remap_table.contents.alphabet_size
        return {
            "handle": t,
            "sigma": int(sigma),
            "len": int(n1),
            "codes": np.ctypeslib.as_array(sa.string, shape=(n1,)).copy(),
            "sa": np.ctypeslib.as_array(sa.array, shape=(n1,)).copy(),
            "c": np.ctypeslib.as_array(tc.c_table, shape=(sigma,)).copy(),
            "o": np.ctypeslib.as_array(tc.o_table, shape=(n1 + 1, sigma)).copy(),
            "table": np.array(list(tc.remap_table.contents.table), dtype=np.int16),
            "ro": np.ctypeslib.as_array(tc.ro_table, shape=(n1 + 1, sigma)).copy() if include_reverse else None,
        }

    def exact_matches(self, handle, codes_pattern: np.ndarray):
        """init/next_bwt_exact_match_iter (bwt.c:164-217): returns (L, R, positions in SA order)."""
        it = RefExactIter()
        m = RefExactMatch()
        pat = np.concatenate([np.asarray(codes_pattern, dtype=np.uint8), np.zeros(1, np.uint8)])
        self.lib.init_bwt_exact_match_iter(C.byref(it), handle, _p(pat, u8p))
        L, R = it.L, it.R
        pos = []
        while self.lib.next_bwt_exact_match_iter(C.byref(it), C.byref(m)):
            pos.append(m.pos)
        return L, R, np.array(pos, dtype=np.uint32)

    def approx_matches(self, handle, codes_pattern: np.ndarray, edits: int):
        """init_bwt_approx_iter / next_bwt_approx_match (bwt.c:302-409): the interval list
        (L, R, match_length, cigar) in report order, and every (position, cigar, match_length)
        the iterator yields."""
        it = RefApproxIter()
        m = RefApproxMatch()
        pat = np.concatenate([np.asarray(codes_pattern, dtype=np.uint8), np.zeros(1, np.uint8)])
        self.lib.init_bwt_approx_iter(C.byref(it), handle, _p(pat, u8p), C.c_int(edits))
        k = it.Ls.used
        L = np.array([it.Ls.data[j] for j in range(k)], dtype=np.uint32)
        R = np.array([it.Rs.data[j] for j in range(k)], dtype=np.uint32)
        ml = np.array([it.match_lengths.data[j] for j in range(k)], dtype=np.uint32)
        cig = [it.cigars.data[j].decode() for j in range(k)]
        hits = []
        while self.lib.next_bwt_approx_match(C.byref(it), C.byref(m)):
            hits.append((int(m.position), m.cigar.decode(), int(m.match_length)))
        self.lib.dealloc_bwt_approx_iter(C.byref(it))
        return L, R, ml, cig, hits

    def free_tables(self, handle):
        self.lib.completely_free_bwt_table(handle)
