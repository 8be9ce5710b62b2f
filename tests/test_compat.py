"""The reference-named drop-in shims (include/stralg_compat.h, libstralg_b200.so).

The GPU tests read like the reference's own tests: tests/stralg/suffix_array_test.c (order,
inverse, LCP, bound searches on "ababacabac"), bwt_test.c ("mississippi" C and O tables,
build_complete_table == tables over a qsort suffix array) and match_test.c (the 8 x 12
string/pattern grid: all four constructors equal, SA-search and BWT-iterator match sets equal
the naive scan).  Expected values come from the golden fixtures and, when the prebuilt
reference library travelled with the snapshot, from calling the same function on the reference.
"""
import ctypes as C
import os
import re

import numpy as np
import pytest

from _oracle import (RefBwtTable, RefExactIter, RefExactMatch, RefRemapTable, RefSuffixArray, bind_stralg_api, u8p,
                     u32p)
from conftest import ROOT

COMPAT_SO = os.path.join(ROOT, "stralg_b200", "lib", "libstralg_b200.so")
STRINGS = ["acacacg", "gacacacag", "acacacag", "acagcaca", "acatgaca", "acgc", "ccgc", "aaaaaaaaa"]
PATTERNS = ["aca", "ac", "ca", "a", "c", "acg", "cg", "g", "cgc", "acgc", "aaa", "aaccaac"]


class SaMatchIter(C.Structure):  # suffix_array.h:74-79
    _fields_ = [("sa", C.POINTER(RefSuffixArray)), ("L", C.c_uint32), ("R", C.c_uint32), ("i", C.c_uint32)]


class SaMatch(C.Structure):
    _fields_ = [("position", C.c_uint32)]


def bind_extra(lib):
    bind_stralg_api(lib)
    lib.init_sa_match_iter.argtypes = [C.POINTER(SaMatchIter), u8p, C.POINTER(RefSuffixArray)]
    lib.init_sa_match_iter.restype = None
    lib.next_sa_match.argtypes = [C.POINTER(SaMatchIter), C.POINTER(SaMatch)]
    lib.next_sa_match.restype = C.c_bool
    lib.init_remap_table.argtypes = [C.POINTER(RefRemapTable), u8p]
    lib.init_remap_table.restype = None
    lib.init_bwt_table.argtypes = [C.POINTER(RefBwtTable), C.POINTER(RefSuffixArray), C.POINTER(RefSuffixArray),
                                   C.POINTER(RefRemapTable)]
    lib.init_bwt_table.restype = None
    lib.dealloc_bwt_table.argtypes = [C.POINTER(RefBwtTable)]
    lib.dealloc_bwt_table.restype = None
    lib.equivalent_bwt_tables.argtypes = [C.POINTER(RefBwtTable), C.POINTER(RefBwtTable)]
    lib.equivalent_bwt_tables.restype = C.c_bool
    lib.identical_suffix_arrays.argtypes = [C.POINTER(RefSuffixArray), C.POINTER(RefSuffixArray)]
    lib.identical_suffix_arrays.restype = C.c_bool
    # index files (suffix_array.h:109-123, remap.h:89-102, bwt.h:337-351, serialise.h:23-33)
    lib.write_suffix_array_fname.argtypes = [C.c_char_p, C.POINTER(RefSuffixArray)]
    lib.write_suffix_array_fname.restype = None
    lib.read_suffix_array_fname.argtypes = [C.c_char_p, u8p]
    lib.read_suffix_array_fname.restype = C.POINTER(RefSuffixArray)
    lib.write_remap_table_fname.argtypes = [C.c_char_p, C.POINTER(RefRemapTable)]
    lib.write_remap_table_fname.restype = None
    lib.read_remap_table_fname.argtypes = [C.c_char_p]
    lib.read_remap_table_fname.restype = C.POINTER(RefRemapTable)
    lib.write_bwt_table_fname.argtypes = [C.c_char_p, C.POINTER(RefBwtTable)]
    lib.write_bwt_table_fname.restype = None
    lib.read_bwt_table_fname.argtypes = [C.c_char_p, C.POINTER(RefSuffixArray), C.POINTER(RefRemapTable)]
    lib.read_bwt_table_fname.restype = C.POINTER(RefBwtTable)
    lib.write_complete_bwt_info_fname.argtypes = [C.c_char_p, C.POINTER(RefBwtTable)]
    lib.write_complete_bwt_info_fname.restype = None
    lib.read_complete_bwt_info_fname.argtypes = [C.c_char_p]
    lib.read_complete_bwt_info_fname.restype = C.POINTER(RefBwtTable)
    return lib


def bind_mapper(lib):
    """The batched read mapper and the multi-record index file (ours, not in the reference library)."""
    lib.bwt_map_fastq_exact.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_char_p),
                                        C.POINTER(C.POINTER(RefBwtTable)), C.c_uint64]
    lib.bwt_map_fastq_exact.restype = C.c_uint64
    lib.bwt_map_fastq.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_char_p),
                                  C.POINTER(C.POINTER(RefBwtTable)), C.c_int, C.c_uint64]
    lib.bwt_map_fastq.restype = C.c_uint64
    lib.write_bwt_tables_file.argtypes = [C.c_char_p, C.c_uint32, C.POINTER(C.c_char_p),
                                          C.POINTER(C.POINTER(RefBwtTable))]
    lib.write_bwt_tables_file.restype = None
    lib.read_bwt_tables_file.argtypes = [C.c_char_p, C.POINTER(C.POINTER(C.c_char_p)),
                                         C.POINTER(C.POINTER(C.POINTER(RefBwtTable)))]
    lib.read_bwt_tables_file.restype = C.c_uint32
    return lib


@pytest.fixture(scope="module")
def compat():
    assert os.path.exists(COMPAT_SO), "libstralg_b200.so missing: run __graft_entry__.build()"
    return bind_extra(C.CDLL(COMPAT_SO))


def cbuf(b: bytes):
    return C.create_string_buffer(b, len(b) + 1)


def declared_compat_symbols():
    text = open(os.path.join(ROOT, "include", "stralg_compat.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"#define[^\n]*", "", text)
    names = set(re.findall(r"\b([a-z_0-9]+)\s*\(", text))
    return sorted(n for n in names if n not in {"defined"})


# ---- CPU -------------------------------------------------------------------------------------------
def test_compat_library_exports_every_declared_symbol(compat):
    names = declared_compat_symbols()
    assert len(names) >= 35
    for name in names:
        assert hasattr(compat, name), f"{name} declared in include/stralg_compat.h but not exported"


def test_struct_layouts_match_reference_abi():
    """Field offsets of the structs callers stack-allocate (suffix_array.h:10-20, bwt.h:36-44,
    bwt.h:168-173, remap.h:9-19) on LP64."""
    assert C.sizeof(RefSuffixArray) == 40 and RefSuffixArray.array.offset == 16 and RefSuffixArray.lcp.offset == 32
    assert C.sizeof(RefBwtTable) == 56 and RefBwtTable.o_indices.offset == 32
    assert C.sizeof(RefExactIter) == 32 and RefExactIter.i.offset == 16 and RefExactIter.R.offset == 24
    assert C.sizeof(RefRemapTable) == 388 and RefRemapTable.rev_table.offset == 260


def test_remap_shims_against_reference(compat, ref, oracle):
    for raw in (b"acagtgtaac", b"mississippi", b"hello, world", bytes(range(1, 100))):
        t = RefRemapTable()
        compat.init_remap_table(C.byref(t), C.cast(cbuf(raw), u8p))
        codes, sigma, table = oracle.remap(raw)
        assert t.alphabet_size == sigma
        assert list(t.table) == table.tolist()
        out = C.create_string_buffer(len(raw) + 1)
        end = compat.remap(C.cast(out, u8p), C.cast(cbuf(raw), u8p), C.byref(t))
        assert end == C.addressof(out) + len(raw) + 1  # pointer past the mapped NUL (remap.c:102-114)
        assert list(out.raw) == codes.tolist()
        if ref is not None:
            t2 = RefRemapTable()
            ref.lib.init_remap_table.argtypes = [C.POINTER(RefRemapTable), u8p]
            ref.lib.init_remap_table(C.byref(t2), C.cast(cbuf(raw), u8p))
            assert bytes(t) == bytes(t2)
    t = RefRemapTable()
    compat.init_remap_table(C.byref(t), C.cast(cbuf(b"acgt"), u8p))
    out = C.create_string_buffer(8)
    assert compat.remap(C.cast(out, u8p), C.cast(cbuf(b"acgx"), u8p), C.byref(t)) is None  # remap.c:80-84


def host_sa(oracle, raw):
    """A struct suffix_array filled from the oracle (no GPU): lets the host-only searches run on CPU."""
    codes, sigma, table = oracle.remap(raw)
    sa = oracle.sa(codes)
    s = RefSuffixArray()
    s.string = codes.ctypes.data_as(u8p)
    s.length = len(codes)
    s.array = sa.ctypes.data_as(u32p)
    return s, codes, sa, table


def test_bound_searches_against_reference(compat, ref, oracle):
    """lower/upper_bound_k, lower/upper_bound_search, the SA match iterator
    (suffix_array.c:90-233; values pinned in tests/stralg/suffix_array_test.c:33-126)."""
    if ref is None:
        pytest.skip("oracle/_ref/libstralg_ref.so not present")
    bind_extra(ref.lib)
    rng = np.random.default_rng(3)
    for raw in [b"ababacabac", b"mississippi"] + [s.encode() for s in STRINGS] + \
            [bytes(rng.choice(list(b"acgt"), 300).astype(np.uint8))]:
        s, codes, sa, table = host_sa(oracle, raw)
        n1 = len(codes)
        sigma = int(codes.max()) + 1
        for k in range(0, 4):
            for a in range(1, sigma):
                for (L, R) in [(0, n1), (1, n1), (n1 // 3, n1), (0, n1 // 2 + 1), (2, 2)]:
                    if L > R:
                        continue
                    assert compat.lower_bound_k(C.byref(s), k, a, L, R) == ref.lib.lower_bound_k(C.byref(s), k, a, L, R)
                    assert compat.upper_bound_k(C.byref(s), k, a, L, R) == ref.lib.upper_bound_k(C.byref(s), k, a, L, R)
        keys = [raw[i:i + m] for i in range(0, min(len(raw), 12)) for m in (1, 2, 3) if i + m <= len(raw)]
        for key in keys:
            kc = table[np.frombuffer(key, dtype=np.uint8)].astype(np.uint8)
            kb = cbuf(bytes(kc))
            assert compat.lower_bound_search(C.byref(s), C.cast(kb, u8p)) == \
                ref.lib.lower_bound_search(C.byref(s), C.cast(kb, u8p))
            assert compat.upper_bound_search(C.byref(s), C.cast(kb, u8p)) == \
                ref.lib.upper_bound_search(C.byref(s), C.cast(kb, u8p))
            got, exp = [], []
            for lib, acc in ((compat, got), (ref.lib, exp)):
                it, m = SaMatchIter(), SaMatch()
                lib.init_sa_match_iter(C.byref(it), C.cast(kb, u8p), C.byref(s))
                while lib.next_sa_match(C.byref(it), C.byref(m)):
                    acc.append(m.position)
            assert got == exp and len(got) >= 1


def table_arrays(tbl):
    tc = tbl.contents
    sigma = tc.remap_table.contents.alphabet_size
    n1 = tc.sa.contents.length
    out = {"string": bytes(np.ctypeslib.as_array(tc.sa.contents.string, shape=(n1,))),
           "sa": np.ctypeslib.as_array(tc.sa.contents.array, shape=(n1,)).copy(),
           "remap": bytes(tc.remap_table.contents),
           "c": np.ctypeslib.as_array(tc.c_table, shape=(sigma,)).copy(),
           "o": np.ctypeslib.as_array(tc.o_table, shape=(n1 + 1, sigma)).copy(),
           "ro": np.ctypeslib.as_array(tc.ro_table, shape=(n1 + 1, sigma)).copy() if tc.ro_table else None}
    return out


def same_tables(a, b):
    return all((a[k] is None and b[k] is None) or np.array_equal(a[k], b[k]) if isinstance(a[k], np.ndarray) or a[k] is None
               else a[k] == b[k] for k in a)


def test_index_files_interoperate_with_the_reference(compat, ref, tmp_path):
    """the on-disk layouts of serialise.c:7-49 / bwt.c:425-503 / suffix_array.c:238-267 /
    remap.c:168-201 (tests/stralg/serialise_test.c round trip).  Host-only: the reference builds the tables,
    each library writes them, the files are byte-identical and each library reads the other's file."""
    if ref is None:
        pytest.skip("oracle/_ref/libstralg_ref.so not present")
    bind_extra(ref.lib)
    for raw, rev in ((b"acgtadtadadfasdfing", False), (b"mississippi", True), (b"a", False)):
        tbl = ref.lib.build_complete_table(C.cast(cbuf(raw), u8p), rev)
        f_ref, f_our = str(tmp_path / "ref.bwt").encode(), str(tmp_path / "our.bwt").encode()
        ref.lib.write_complete_bwt_info_fname(f_ref, tbl)
        compat.write_complete_bwt_info_fname(f_our, tbl)  # same struct layouts: the shim writes the reference's table
        assert open(f_ref, "rb").read() == open(f_our, "rb").read()
        exp = table_arrays(tbl)
        got_our = compat.read_complete_bwt_info_fname(f_ref)
        got_ref = ref.lib.read_complete_bwt_info_fname(f_our)
        assert same_tables(exp, table_arrays(got_our)) and same_tables(exp, table_arrays(got_ref))
        assert bool(got_our.contents.ro_table) == rev
        # the O(a, i) macro path of a table read from a file: o_indices[i][a]
        n1, sigma = len(raw) + 1, exp["c"].shape[0]
        for i in (0, n1 // 2, n1):
            assert np.ctypeslib.as_array(got_our.contents.o_indices[i], shape=(sigma,)).tolist() == exp["o"][i].tolist()
        # the piecewise files
        fs, fr, fb = (str(tmp_path / n).encode() for n in ("x.sa", "x.remap", "x.tbl"))
        compat.write_suffix_array_fname(fs, tbl.contents.sa)
        compat.write_remap_table_fname(fr, tbl.contents.remap_table)
        compat.write_bwt_table_fname(fb, tbl)
        sa2 = ref.lib.read_suffix_array_fname(fs, tbl.contents.sa.contents.string)
        rt2 = ref.lib.read_remap_table_fname(fr)
        tb2 = ref.lib.read_bwt_table_fname(fb, sa2, rt2)
        assert same_tables(exp, table_arrays(tb2))
        sa3 = compat.read_suffix_array_fname(fs, tbl.contents.sa.contents.string)
        rt3 = compat.read_remap_table_fname(fr)
        tb3 = compat.read_bwt_table_fname(fb, sa3, rt3)
        assert same_tables(exp, table_arrays(tb3))


# ---- GPU -------------------------------------------------------------------------------------------
def naive_positions(text: bytes, pat: bytes):
    return [i for i in range(len(text) - len(pat) + 1) if text[i:i + len(pat)] == pat]


@pytest.mark.gpu
def test_suffix_array_test_c(compat, engine, golden):
    """tests/stralg/suffix_array_test.c: "ababacabac" through all four constructors."""
    raw = b"ababacabac"
    exp_sa = golden["ababacabac/sa"]
    t = RefRemapTable()
    compat.init_remap_table(C.byref(t), C.cast(cbuf(raw), u8p))
    remapped = C.create_string_buffer(len(raw) + 1)
    compat.remap(C.cast(remapped, u8p), C.cast(cbuf(raw), u8p), C.byref(t))
    for ctor, args in ((compat.qsort_sa_construction, ()), (compat.skew_sa_construction, ()),
                       (compat.sa_is_construction, (t.alphabet_size,)),
                       (compat.sa_is_mem_construction, (t.alphabet_size,))):
        sa = ctor(C.cast(remapped, u8p), *args)
        n1 = sa.contents.length
        assert n1 == 11
        arr = np.ctypeslib.as_array(sa.contents.array, shape=(n1,))
        assert arr.tolist() == exp_sa.tolist() == [10, 0, 6, 2, 8, 4, 1, 7, 3, 9, 5]
        # test_order (suffix_array_test.c:11-17): adjacent suffixes strictly increasing
        sufs = [remapped.raw[i:].split(b"\0")[0] for i in arr]
        assert all(a < b for a, b in zip(sufs, sufs[1:]))
        compat.compute_lcp(sa)  # also fills inverse (suffix_array.c:70)
        inv = np.ctypeslib.as_array(sa.contents.inverse, shape=(n1,))
        lcp = np.ctypeslib.as_array(sa.contents.lcp, shape=(n1,))
        assert all(inv[arr[i]] == i for i in range(n1)) and all(arr[inv[i]] == i for i in range(n1))
        assert lcp[0] == 0
        for i in range(1, n1):  # brute-force LCP (suffix_array_test.c:138-158)
            a, b = sufs[i - 1], sufs[i]
            l = 0
            while l < min(len(a), len(b)) and a[l] == b[l]:
                l += 1
            assert lcp[i] == l
        assert lcp.tolist() == golden["ababacabac/lcp"].tolist()
        compat.compute_lcp(sa)  # idempotent
        compat.free_suffix_array(sa)


@pytest.mark.gpu
def test_bwt_test_c(compat, engine, golden):
    """tests/stralg/bwt_test.c: "mississippi" C and O tables; build_complete_table equals the
    table over a qsort suffix array (bwt_test.c:187-189)."""
    raw = b"mississippi"
    tbl = compat.build_complete_table(C.cast(cbuf(raw), u8p), False)
    tc = tbl.contents
    sigma = tc.remap_table.contents.alphabet_size
    n1 = tc.sa.contents.length
    assert sigma == 5 and n1 == 12
    assert np.ctypeslib.as_array(tc.sa.contents.string, shape=(n1,)).tolist() == [2, 1, 4, 4, 1, 4, 4, 1, 3, 3, 1, 0]
    assert np.ctypeslib.as_array(tc.c_table, shape=(sigma,)).tolist() == [0, 1, 5, 6, 8]
    o = np.ctypeslib.as_array(tc.o_table, shape=(n1 + 1, sigma))
    assert np.array_equal(o, golden["mississippi/o"])
    # the O(a, i) macro path: o_indices[i][a]
    for i in (0, 5, 12):
        row = np.ctypeslib.as_array(tc.o_indices[i], shape=(sigma,))
        assert row.tolist() == golden["mississippi/o"][i].tolist()
    # second table over a qsort suffix array of the same remapped string
    sa2 = compat.qsort_sa_construction(tc.sa.contents.string)
    tbl2 = RefBwtTable()
    compat.init_bwt_table(C.byref(tbl2), sa2, None, tc.remap_table)
    assert compat.equivalent_bwt_tables(tbl, C.byref(tbl2))
    compat.dealloc_bwt_table(C.byref(tbl2))
    compat.free_suffix_array(sa2)
    # reverse table requested: RO equals the O table of the reversed text
    tblr = compat.build_complete_table(C.cast(cbuf(raw), u8p), True)
    assert bool(tblr.contents.ro_table)
    ro = np.ctypeslib.as_array(tblr.contents.ro_table, shape=(n1 + 1, sigma))
    rev = compat.build_complete_table(C.cast(cbuf(raw[::-1]), u8p), False)
    assert np.array_equal(ro, np.ctypeslib.as_array(rev.contents.o_table, shape=(n1 + 1, sigma)))
    for t in (tbl, tblr, rev):
        compat.completely_free_bwt_table(t)


@pytest.mark.gpu
def test_match_test_c(compat, engine, golden):
    """tests/stralg/match_test.c:682-696 grid: constructors agree, SA search and BWT iterator
    match sets equal the naive scan (sorted before comparing, match_test.c:608)."""
    for si, s in enumerate(STRINGS):
        raw = s.encode()
        t = RefRemapTable()
        compat.init_remap_table(C.byref(t), C.cast(cbuf(raw), u8p))
        remapped = C.create_string_buffer(len(raw) + 1)
        compat.remap(C.cast(remapped, u8p), C.cast(cbuf(raw), u8p), C.byref(t))
        sas = [compat.qsort_sa_construction(C.cast(remapped, u8p)), compat.skew_sa_construction(C.cast(remapped, u8p)),
               compat.sa_is_construction(C.cast(remapped, u8p), t.alphabet_size),
               compat.sa_is_mem_construction(C.cast(remapped, u8p), t.alphabet_size)]
        for other in sas[1:]:
            assert compat.identical_suffix_arrays(sas[0], other)
        arr = np.ctypeslib.as_array(sas[0].contents.array, shape=(len(raw) + 1,))
        assert np.array_equal(arr, golden[f"grid{si}/sa"])
        tbl = RefBwtTable()
        compat.init_bwt_table(C.byref(tbl), sas[0], None, C.byref(t))
        for pi, p in enumerate(PATTERNS):
            pm = C.create_string_buffer(len(p) + 1)
            if not compat.remap(C.cast(pm, u8p), C.cast(cbuf(p.encode()), u8p), C.byref(t)):
                continue  # letters not in the text (match_test.c:635)
            naive = naive_positions(raw, p.encode())
            it, m = SaMatchIter(), SaMatch()
            got = []
            compat.init_sa_match_iter(C.byref(it), C.cast(pm, u8p), sas[2])
            while compat.next_sa_match(C.byref(it), C.byref(m)):
                got.append(m.position)
            assert sorted(got) == naive, (s, p)
            bit, bm = RefExactIter(), RefExactMatch()
            compat.init_bwt_exact_match_iter(C.byref(bit), C.byref(tbl), C.cast(pm, u8p))
            got = []
            while compat.next_bwt_exact_match_iter(C.byref(bit), C.byref(bm)):
                got.append(bm.pos)
            assert sorted(got) == naive, (s, p)
            key = f"grid{si}/pat{pi}/pos"
            assert got == golden[key].tolist()  # same SA order as the reference iterator
        compat.dealloc_bwt_table(C.byref(tbl))
        for sa in sas:
            compat.free_suffix_array(sa)


@pytest.mark.gpu
def test_serialise_test_c(compat, engine, ref, tmp_path):
    """tests/stralg/serialise_test.c:13-43: build_complete_table -> file -> read back -> equivalent tables, with the
    tables built on the GPU; the table read from the file then serves the exact iterator (its device index is
    rebuilt from the string on first use)."""
    raw = b"acgtadtadadfasdfing"
    tbl = compat.build_complete_table(C.cast(cbuf(raw), u8p), False)
    fname = str(tmp_path / "index.bwttables").encode()
    compat.write_complete_bwt_info_fname(fname, tbl)
    back = compat.read_complete_bwt_info_fname(fname)
    assert compat.equivalent_bwt_tables(tbl, back)
    assert same_tables(table_arrays(tbl), table_arrays(back))
    if ref is not None:
        bind_extra(ref.lib)
        theirs = ref.lib.build_complete_table(C.cast(cbuf(raw), u8p), False)
        fref = str(tmp_path / "ref.bwttables").encode()
        ref.lib.write_complete_bwt_info_fname(fref, theirs)
        assert open(fref, "rb").read() == open(fname, "rb").read()  # a GPU-built index file == the reference's
    for pat in (b"ad", b"tad", b"a", b"ing", b"gg"):
        pm = C.create_string_buffer(len(pat) + 1)
        assert compat.remap(C.cast(pm, u8p), C.cast(cbuf(pat), u8p), back.contents.remap_table)
        bit, bm = RefExactIter(), RefExactMatch()
        compat.init_bwt_exact_match_iter(C.byref(bit), back, C.cast(pm, u8p))
        got = []
        while compat.next_bwt_exact_match_iter(C.byref(bit), C.byref(bm)):
            got.append(bm.pos)
        assert sorted(got) == naive_positions(raw, pat)
    compat.completely_free_bwt_table(back)
    compat.completely_free_bwt_table(tbl)


def read_fasta(path):
    recs, name, seq = [], None, []
    for ln in open(path):
        ln = ln.strip()
        if ln.startswith(">"):
            if name is not None:
                recs.append((name, "".join(seq)))
            name, seq = ln[1:].strip(), []
        elif ln:
            seq.append(ln)
    recs.append((name, "".join(seq)))
    return recs


@pytest.mark.gpu
@pytest.mark.parametrize("batch_reads", [0, 7])
def test_readmapper_exact_matches_reference_tool(compat, engine, tmp_path, batch_reads):
    """`bwt_readmapper -p` + `bwt_readmapper -d 0` (tools/readmappers/bwt_readmapper/
    bwt_readmapper.c:16-67, 128-161) against the golden files the reference tool produced
    (tests/golden/make_readmapper_golden.py): the index file is byte-identical (sha256), the SAM output is
    byte-identical, with the tables built and the reads searched on the GPU."""
    import hashlib
    import json
    gdir = os.path.join(ROOT, "tests", "golden", "readmapper")
    meta = json.load(open(os.path.join(gdir, "meta.json")))
    bind_mapper(compat)
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    recs = read_fasta(os.path.join(gdir, "ref.fa"))
    assert [r[0] for r in recs] == ["chrA", "chrB"]
    # -p: records in reverse FASTA order (bioinf/fasta.c:131), tables with the reverse O table (bwt_readmapper.c:58)
    order = recs[::-1]
    tbls = [compat.build_complete_table(C.cast(cbuf(seq.encode()), u8p), True) for _, seq in order]
    names = (C.c_char_p * len(order))(*[n.encode() for n, _ in order])
    tarr = (C.POINTER(RefBwtTable) * len(order))(*tbls)
    fname = str(tmp_path / "ref.fa.bwttables").encode()
    compat.write_bwt_tables_file(fname, len(order), names, tarr)
    blob = open(fname, "rb").read()
    assert len(blob) == meta["bwttables_bytes"]
    assert hashlib.sha256(blob).hexdigest() == meta["bwttables_sha256"]
    for t in tbls:
        compat.completely_free_bwt_table(t)
    # -d 0: read the file back, map in the reverse of the file order (bwt_readmapper.c:107)
    rnames, rtabs = C.POINTER(C.c_char_p)(), C.POINTER(C.POINTER(RefBwtTable))()
    nrec = compat.read_bwt_tables_file(fname, C.byref(rnames), C.byref(rtabs))
    assert nrec == 2 and [rnames[i] for i in range(nrec)] == [b"chrB", b"chrA"]
    mnames = (C.c_char_p * nrec)(*[rnames[i] for i in reversed(range(nrec))])
    mtabs = (C.POINTER(RefBwtTable) * nrec)(*[rtabs[i] for i in reversed(range(nrec))])
    fq = libc.fopen(os.path.join(gdir, "reads.fq").encode(), b"r")
    out_path = str(tmp_path / "out.sam").encode()
    out = libc.fopen(out_path, b"w")
    nlines = compat.bwt_map_fastq_exact(fq, out, nrec, mnames, mtabs, batch_reads)
    libc.fclose(fq)
    libc.fclose(out)
    assert nlines == meta["sam_lines"]
    assert open(out_path, "rb").read() == open(os.path.join(gdir, "expected.sam"), "rb").read()
    for i in range(nrec):
        compat.completely_free_bwt_table(rtabs[i])


@pytest.mark.gpu
@pytest.mark.parametrize("edits,batch_reads", [(1, 0), (1, 5), (2, 0)])
def test_readmapper_approx_matches_reference_tool(compat, engine, tmp_path, edits, batch_reads):
    """SURVEY 8f rank 4 end to end: `bwt_readmapper -d 1` / `-d 2` (map_read, bwt_readmapper.c:128-161,
    through init_bwt_approx_iter / next_bwt_approx_match) against the SAM files the reference tool wrote:
    every (read, record, position, CIGAR) line, in the tool's order, byte for byte."""
    import json
    gdir = os.path.join(ROOT, "tests", "golden", "readmapper")
    meta = json.load(open(os.path.join(gdir, "meta.json")))
    bind_mapper(compat)
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    recs = read_fasta(os.path.join(gdir, "ref.fa"))
    # the tool maps against the records in FASTA order (two prepends cancel, bwt_readmapper.c:48-61, 107)
    tbls = [compat.build_complete_table(C.cast(cbuf(seq.encode()), u8p), True) for _, seq in recs]
    names = (C.c_char_p * len(recs))(*[n.encode() for n, _ in recs])
    tarr = (C.POINTER(RefBwtTable) * len(recs))(*tbls)
    fq = libc.fopen(os.path.join(gdir, f"reads_d{edits}.fq").encode(), b"r")
    out_path = str(tmp_path / "out.sam").encode()
    out = libc.fopen(out_path, b"w")
    nlines = compat.bwt_map_fastq(fq, out, len(recs), names, tarr, edits, batch_reads)
    libc.fclose(fq)
    libc.fclose(out)
    assert nlines == meta[f"d{edits}_sam_lines"]
    assert open(out_path, "rb").read() == open(os.path.join(gdir, f"expected_d{edits}.sam"), "rb").read()
    for t in tbls:
        compat.completely_free_bwt_table(t)


@pytest.mark.gpu
def test_compat_approx_iterator_matches_reference(compat, engine, ref):
    """init_bwt_approx_iter / next_bwt_approx_match / dealloc_bwt_approx_iter with the reference's struct
    layouts: same (position, cigar, match_length) sequence as the unmodified reference on mississippi
    (the string of tests/stralg/bwt_test.c) and a random DNA text, with and without the RO table."""
    from _oracle import RefApproxIter, RefApproxMatch
    if ref is None:
        pytest.skip("oracle/_ref/libstralg_ref.so did not travel with this snapshot")
    compat.init_bwt_approx_iter.argtypes = [C.POINTER(RefApproxIter), C.POINTER(RefBwtTable), u8p, C.c_int]
    compat.init_bwt_approx_iter.restype = None
    compat.next_bwt_approx_match.argtypes = [C.POINTER(RefApproxIter), C.POINTER(RefApproxMatch)]
    compat.next_bwt_approx_match.restype = C.c_bool
    compat.dealloc_bwt_approx_iter.argtypes = [C.POINTER(RefApproxIter)]
    compat.dealloc_bwt_approx_iter.restype = None
    rng = np.random.default_rng(5)
    for raw in (b"mississippi", bytes(rng.choice(list(b"acgt"), 900).tolist())):
        for with_ro in (True, False):
            tbl = compat.build_complete_table(C.cast(cbuf(raw), u8p), with_ro)
            rt = ref.tables(raw, include_reverse=with_ro)
            pats = [b"ssi", b"is", b"mississippi", b"ppi"] if raw == b"mississippi" else \
                   [raw[s:s + m] for s, m in ((0, 5), (100, 9), (500, 12), (880, 7))]
            for p in pats:
                pm = C.create_string_buffer(len(p) + 1)
                assert compat.remap(C.cast(pm, u8p), C.cast(cbuf(p), u8p), tbl.contents.remap_table)
                codes = np.frombuffer(pm.raw[:len(p)], dtype=np.uint8)
                for d in (0, 1, 2):
                    it, m = RefApproxIter(), RefApproxMatch()
                    compat.init_bwt_approx_iter(C.byref(it), tbl, C.cast(pm, u8p), d)
                    got = []
                    while compat.next_bwt_approx_match(C.byref(it), C.byref(m)):
                        got.append((int(m.position), m.cigar.decode(), int(m.match_length)))
                    compat.dealloc_bwt_approx_iter(C.byref(it))
                    assert got == ref.approx_matches(rt["handle"], codes, d)[4], (raw[:12], p, d, with_ro)
            ref.free_tables(rt["handle"])
            compat.completely_free_bwt_table(tbl)


@pytest.mark.gpu
def test_compat_large_text_files(compat, engine, golden):
    """The two data-file runs of tests/stralg/CMakeLists.txt:19-30: "the" in modest-proposal.txt and
    "ababaaba" in repetitive-string.txt."""
    for name, pat in (("modest", b"the"), ("repetitive", b"ababaaba")):
        raw = bytes(golden[f"{name}/raw"])
        tbl = compat.build_complete_table(C.cast(cbuf(raw), u8p), False)
        tc = tbl.contents
        n1 = tc.sa.contents.length
        assert np.array_equal(np.ctypeslib.as_array(tc.sa.contents.array, shape=(n1,)), golden[f"{name}/sa"])
        pm = C.create_string_buffer(len(pat) + 1)
        assert compat.remap(C.cast(pm, u8p), C.cast(cbuf(pat), u8p), tc.remap_table)
        bit, bm = RefExactIter(), RefExactMatch()
        compat.init_bwt_exact_match_iter(C.byref(bit), tbl, C.cast(pm, u8p))
        got = []
        while compat.next_bwt_exact_match_iter(C.byref(bit), C.byref(bm)):
            got.append(bm.pos)
        assert sorted(got) == naive_positions(raw, pat) and len(got) > 0
        compat.completely_free_bwt_table(tbl)
