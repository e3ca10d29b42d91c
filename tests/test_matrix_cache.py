"""Binary matrix cache: CSR and BS-CSR round trips through the C ABI, rejection of truncated,
altered and foreign files.  CPU only."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def csr(gen):
    x, y, v = gen.create_sparse_matrix(3000, 1024, 20, "gamma", seed=2)
    return gen.csr_from_coo(x, 3000).astype(np.uint64), y, v.astype(np.float32), x


def test_csr_round_trip(tks, csr, tmp_path):
    ptr, idx, val, _ = csr
    f = tmp_path / "m.tkscsr"
    tks.capi.cache_write_csr(f, ptr, idx, val, 1024)
    rows, cols, p2, i2, v2 = tks.capi.cache_read_csr(f)
    assert (rows, cols) == (3000, 1024)
    assert np.array_equal(p2, ptr) and np.array_equal(i2, idx) and np.array_equal(v2.view(np.uint32), val.view(np.uint32))
    assert not (tmp_path / "m.tkscsr.tmp").exists()


def test_empty_matrix_round_trip(tks, tmp_path):
    f = tmp_path / "e.tkscsr"
    tks.capi.cache_write_csr(f, np.zeros(5, np.uint64), np.zeros(0, np.uint32), np.zeros(0, np.float32), 7)
    rows, cols, p2, i2, v2 = tks.capi.cache_read_csr(f)
    assert (rows, cols, i2.size) == (4, 7, 0)


def test_bscsr_round_trip_equals_fresh_packing(tks, orc, csr, tmp_path):
    ptr, idx, val, x = csr
    val32 = tks.capi.fixed32_from_double_np(val.astype(np.float64))
    for W, P in ((20, 32), (32, 8)):
        packets, ppp, first, npp = tks.capi.pack_bscsr(x, idx, val32, 3000, P, W)
        f = tmp_path / f"m_w{W}.tksbs"
        tks.capi.cache_write_bscsr(f, 3000, 1024, W, packets, ppp, first, npp)
        rows, cols, W2, pk2, ppp2, first2, npp2 = tks.capi.cache_read_bscsr(f)
        assert (rows, cols, W2) == (3000, 1024, W)
        assert np.array_equal(pk2, packets) and np.array_equal(ppp2, ppp) and np.array_equal(first2, first)
        assert np.array_equal(npp2, npp)


def test_corruption_is_detected(tks, csr, tmp_path):
    ptr, idx, val, _ = csr
    f = tmp_path / "m.tkscsr"
    tks.capi.cache_write_csr(f, ptr, idx, val, 1024)
    raw = bytearray(f.read_bytes())
    # one flipped payload bit
    bad = bytearray(raw); bad[len(bad) // 2] ^= 0x10
    (tmp_path / "flip.tkscsr").write_bytes(bad)
    with pytest.raises(tks.capi.TksError, match="checksum"):
        tks.capi.cache_read_csr(tmp_path / "flip.tkscsr")
    # truncated
    (tmp_path / "cut.tkscsr").write_bytes(raw[:-100])
    with pytest.raises(tks.capi.TksError, match="truncated"):
        tks.capi.cache_read_csr(tmp_path / "cut.tkscsr")
    # not a cache file / the other kind / missing
    (tmp_path / "txt.tkscsr").write_bytes(b"%%MatrixMarket matrix coordinate real general\n" * 4)
    with pytest.raises(tks.capi.TksError, match="not a TKSMAT01"):
        tks.capi.cache_read_csr(tmp_path / "txt.tkscsr")
    with pytest.raises(tks.capi.TksError, match="different kind"):
        tks.capi.cache_read_bscsr(f)
    with pytest.raises(tks.capi.TksError, match="not found"):
        tks.capi.cache_read_csr(tmp_path / "nope.tkscsr")
    # the writer refuses an inconsistent CSR
    with pytest.raises(tks.capi.TksError, match="ptr"):
        tks.capi.cache_write_csr(tmp_path / "x", ptr[:-1], idx, val, 1024)


def test_csr_cache_carries_a_tag_of_its_source(tks, tmp_path):
    """The host executable only trusts a cache made from the same Matrix-Market file (path, size, mtime) with the same
    -z / -v flags (main_b200.cpp): the tag travels in the header."""
    import ctypes as C
    import os
    L = tks.capi.lib()
    L.tks_cache_source_tag.argtypes = [C.c_char_p, C.c_int, C.c_int]
    L.tks_cache_source_tag.restype = C.c_uint64
    L.tks_cache_write_csr_tagged.argtypes = [C.c_char_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    L.tks_cache_read_csr_tagged.argtypes = [C.c_char_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64),
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
    src = tmp_path / "m.mtx"
    src.write_text("%%MatrixMarket matrix coordinate real general\n%\n2 2 2\n1 1 0.5\n2 2 0.25\n")
    t0 = L.tks_cache_source_tag(str(src).encode(), 0, 0)
    assert t0 != 0 and t0 == L.tks_cache_source_tag(str(src).encode(), 0, 0)
    assert t0 != L.tks_cache_source_tag(str(src).encode(), 1, 0) and t0 != L.tks_cache_source_tag(str(src).encode(), 0, 1)
    ptr = np.array([0, 1, 2], np.uint64)
    idx = np.array([0, 1], np.uint32)
    val = np.array([0.5, 0.25], np.float32)
    cache = str(tmp_path / "m.tkscsr").encode()
    assert L.tks_cache_write_csr_tagged(cache, 2, 2, 2, ptr.ctypes.data, idx.ctypes.data, val.ctypes.data, t0) == 0
    rows, cols, nnz, tag = C.c_uint64(), C.c_uint32(), C.c_uint64(), C.c_uint64()
    assert L.tks_cache_read_csr_tagged(cache, C.byref(rows), C.byref(cols), C.byref(nnz), None, None, None, C.byref(tag)) == 0
    assert (rows.value, cols.value, nnz.value, tag.value) == (2, 2, 2, t0)
    # the file changes (content and mtime): its tag no longer matches the cache's
    src.write_text("%%MatrixMarket matrix coordinate real general\n%\n2 2 2\n1 1 0.75\n2 2 0.25\n")
    os.utime(src, ns=(1, 1))
    assert L.tks_cache_source_tag(str(src).encode(), 0, 0) != t0
    # untagged writes read back as tag 0
    tks.capi.cache_write_csr(tmp_path / "u.tkscsr", ptr, idx, val, 2)
    assert L.tks_cache_read_csr_tagged(str(tmp_path / "u.tkscsr").encode(), C.byref(rows), C.byref(cols), C.byref(nnz), None, None, None, C.byref(tag)) == 0
    assert tag.value == 0
