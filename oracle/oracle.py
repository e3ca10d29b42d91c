"""ctypes/numpy face of the CPU oracle (oracle/oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / ``--impl reference`` legs, never by the product
package.  Each function cites the reference file:line it restates (paths are
relative to the reference checkout).

The float64 stand-in for ``sparse_dot_topn.awesome_cossim_topn`` (test_cpu.py:104)
is "parity unpinned": that dependency is absent from the image and un-pinned
in the reference (README.md:49), so its published semantics are restated
(CSR x dense column in float64, entries <= lower_bound dropped, top-n).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "liboracle.so"
_REF_GOLD_PATH = _HERE / "_ref" / "libref_gold.so"
_REF_FPGA_PATH = _HERE / "_ref" / "libref_fpga.so"

_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")


def build(ref: bool = True) -> None:
    """Compile oracle.c (always) and, where /root/reference exists, oracle/_ref."""
    subprocess.run(["make", "-s", "-C", str(_HERE), "_build/liboracle.so"], check=True)
    if ref and Path(os.environ.get("TKS_REFERENCE_ROOT", "/root/reference")).is_dir():
        subprocess.run(["make", "-s", "-C", str(_HERE), "ref"], check=True)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            build(ref=False)
        L = C.CDLL(str(_LIB_PATH))
        L.orc_gold_topk_f32.argtypes = [_u32p, _u32p, _f32p, C.c_uint64, _f32p, C.c_int, _u32p, _f32p]
        L.orc_sort_tuples_f32.argtypes = [C.c_uint32, _u32p, _f32p]
        L.orc_sort_tuples_u32.argtypes = [C.c_uint32, _u32p, _u32p]
        L.orc_spmv_f32.argtypes = [_u32p, _u32p, _f32p, C.c_uint64, _f32p, _f32p, C.c_uint32]
        L.orc_fx32_from_double.argtypes = [C.c_double]
        L.orc_fx32_from_double.restype = C.c_uint32
        L.orc_fxW_from_fx32.argtypes = [C.c_uint32, C.c_int]
        L.orc_fxW_from_fx32.restype = C.c_uint32
        L.orc_packet_size.argtypes = [C.c_int]
        L.orc_packet_size.restype = C.c_int
        L.orc_partition.argtypes = [_u32p, C.c_uint64, C.c_uint32, C.c_int, C.c_int, _u64p, _u32p, _u32p, _u64p]
        L.orc_partition.restype = C.c_int
        L.orc_pack_partition.argtypes = [_u32p, _u32p, _u32p, C.c_uint64, C.c_uint32, C.c_int, _u64p]
        L.orc_pack_query.argtypes = [_u32p, C.c_uint32, C.c_int, _u64p]
        L.orc_bscsr_partition.argtypes = [_u64p, C.c_uint64, _u32p, C.c_uint32, C.c_int, C.c_int, C.c_int,
                                          _u32p, _u32p]
        L.orc_bscsr_partition_ex.argtypes = [_u64p, C.c_uint64, _u32p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int,
                                             _u32p, _u32p]
        L.orc_read_result.argtypes = [C.c_int, C.c_int, C.c_int, _u32p, _u32p, _u32p, _u32p, _u32p]
        L.orc_read_result.restype = C.c_uint32
        L.orc_gold_topk_fx32.argtypes = [_u32p, _u32p, _u32p, C.c_uint64, _u32p, C.c_int, _u32p, _u32p]
        _lib = L
    return _lib


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


# --------------------------------------------------------------------------
# float path
# --------------------------------------------------------------------------

def gold_topk_f32(row, col, val, vec, k, sort=True):
    """gold_algorithms.hpp:188-246 (+ evaluation_utils.hpp:40-62 when sort)."""
    row, col, val, vec = _c(row, np.uint32), _c(col, np.uint32), _c(val, np.float32), _c(vec, np.float32)
    idx = np.zeros(k, np.uint32)
    out = np.zeros(k, np.float32)
    lib().orc_gold_topk_f32(row, col, val, row.size, vec, k, idx, out)
    if sort:
        lib().orc_sort_tuples_f32(k, idx, out)
    return idx, out


def half_round(a):
    """float_to_half then half_to_float (host_spmv_topk_csr_gpu.cu:151-153, 254; IEEE round to nearest even):
    what the reference's half-precision GPU mode does to every matrix value and to the query."""
    return np.asarray(a, np.float32).astype(np.float16).astype(np.float32)


def bf16_round(a):
    """float -> bfloat16 (round to nearest even, what __float2bfloat16_rn does) -> float: keep the upper 16 bits."""
    b = np.ascontiguousarray(a, np.float32).view(np.uint32).astype(np.uint64)
    b = (b + np.uint64(0x7FFF) + ((b >> np.uint64(16)) & np.uint64(1))) & np.uint64(0xFFFF0000)
    return b.astype(np.uint32).view(np.float32)


def gold_topk_bf16(row, col, val, vec, k, sort=True):
    """The gold on bfloat16-rounded inputs with fp32 accumulation: the statement of the engine's TKS_VALUE_BF16 mode
    (not a mode of the reference).  A bf16 x bf16 product is exact in fp32 (8 + 8 significand bits)."""
    return gold_topk_f32(row, col, bf16_round(val), bf16_round(vec), k, sort)


def gold_topk_f16(row, col, val, vec, k, sort=True):
    """The gold on half-rounded inputs with fp32 accumulation: the statement of the engine's TKS_VALUE_FP16 mode.
    (The reference accumulates in half through cuSPARSE CUDA_R_16F / light_spmv<half>, which loses ~3 digits; the
    engine keeps the reference's storage format and rounding of the INPUTS and is at least as accurate after.)"""
    return gold_topk_f32(row, col, half_round(val), half_round(vec), k, sort)


def spmv_f32(row, col, val, vec, num_rows):
    """Sequential fp32 product vector, same order as the gold (gold_algorithms.hpp:203-213)."""
    row, col, val, vec = _c(row, np.uint32), _c(col, np.uint32), _c(val, np.float32), _c(vec, np.float32)
    out = np.zeros(num_rows, np.float32)
    lib().orc_spmv_f32(row, col, val, row.size, vec, out, num_rows)
    return out


def topk_total_order(scores, k, tie="lower", positive_only=False):
    """Top-k of a dense score vector under the stated total order:
    score descending, ties -> lower index first (``tie='lower'``, the north-star
    contract) or higher index first (``tie='higher'``, evaluation_utils.hpp:54-55)."""
    scores = np.asarray(scores)
    n = scores.size
    idx = np.arange(n, dtype=np.int64)
    if positive_only:
        keep = scores > 0
        idx = idx[keep]
    s = scores[idx]
    sec = idx if tie == "lower" else -idx
    order = np.lexsort((sec, -s.astype(np.float64)))[:k]
    return idx[order].astype(np.uint32), s[order]


def f64_topk(ptr, col, val, vec, k, lower_bound=0.0, tie="lower"):
    """Stand-in for test_cpu.py:91-105: float64 CSR x dense query, entries
    <= lower_bound dropped (awesome_cossim_topn's threshold), global top-k."""
    import scipy.sparse as sp

    ptr = np.asarray(ptr, dtype=np.int64)
    n = ptr.size - 1
    a = sp.csr_matrix((np.asarray(val, np.float64), np.asarray(col, np.int64), ptr),
                      shape=(n, int(np.asarray(vec).size)))
    y = a @ np.asarray(vec, np.float64)
    y = np.where(y > lower_bound, y, 0.0)
    return topk_total_order(y, k, tie=tie, positive_only=True)


# --------------------------------------------------------------------------
# fixed-point path
# --------------------------------------------------------------------------

def packet_size(W):
    return int(lib().orc_packet_size(W))


def fx32_from_double(a):
    """double -> raw ap_ufixed<32,1,AP_TRN_ZERO> (utils.hpp:401, :242)."""
    a = np.asarray(a, np.float64)
    s = np.floor(np.where(a > 0, a, 0.0) * 2147483648.0)
    return np.mod(s, 4294967296.0).astype(np.uint64).astype(np.uint32)


def fxW_from_fx32(raw32, W):
    """fpga_utils.hpp:336-338: to_float() (RNE) then truncate to ap_ufixed<W,1>."""
    raw32 = np.asarray(raw32, np.uint32)
    f = (raw32.astype(np.float64) / 2147483648.0).astype(np.float32)
    s = np.floor(f.astype(np.float64) * float(1 << (W - 1))).astype(np.uint64)
    m = np.uint64(0xFFFFFFFF if W == 32 else (1 << W) - 1)
    return (s & m).astype(np.uint32)


def query_fx32_from_f32(vec_f32):
    """create_sample_vector<real_type_inout> tail (utils.hpp:258-266): the
    normalised float value is cast to ap_ufixed<32,1> by truncation."""
    return fx32_from_double(np.asarray(vec_f32, np.float32).astype(np.float64))


def pack_bscsr(row, col, val32, num_rows, P=32, W=20):
    """host_spmv_bscsr.cpp:112-121,133-248.  Returns dict with per-partition
    packet arrays (uint64 [npk,8]), first_row, last_row, nnz_start."""
    L = lib()
    row, col, val32 = _c(row, np.uint32), _c(col, np.uint32), _c(val32, np.uint32)
    B = packet_size(W)
    nnz_start = np.zeros(P + 1, np.uint64)
    first_row = np.zeros(P, np.uint32)
    last_row = np.zeros(P, np.uint32)
    npk = np.zeros(P, np.uint64)
    rc = L.orc_partition(row, row.size, num_rows, P, B, nnz_start, first_row, last_row, npk)
    if rc != 0:
        raise ValueError(f"orc_partition failed rc={rc} (empty partition or unsorted rows)")
    packets = []
    for p in range(P):
        s, e = int(nnz_start[p]), int(nnz_start[p + 1])
        out = np.zeros((int(npk[p]), 8), np.uint64)
        prev_last = 0 if p == 0 else int(last_row[p - 1])
        L.orc_pack_partition(row[s:e].copy(), col[s:e].copy(), val32[s:e].copy(), e - s, prev_last, W,
                             out.reshape(-1))
        packets.append(out)
    return dict(packets=packets, first_row=first_row, last_row=last_row, nnz_start=nnz_start,
                num_packets=npk, B=B, W=W, P=P)


def pack_query(vec32, W):
    vec32 = _c(vec32, np.uint32)
    B = packet_size(W)
    nblk = (vec32.size + B - 1) // B
    out = np.zeros((nblk, 8), np.uint64)
    lib().orc_pack_query(vec32, vec32.size, W, out.reshape(-1))
    return out


def bscsr_kernel(packed, vec32, Kp=8, LFR=4, drift_free=False):
    """spmv_bscsr_top_k_multicore.{hpp,cpp}: all partitions, reference result layout.
    Returns (idx_words[P,Kp,16], val_words[P,Kp,16]).  drift_free=True is NOT the reference: it is the
    engine's stated repair of the row-counter drift (see orc_bscsr_partition_ex)."""
    L = lib()
    vec32 = _c(vec32, np.uint32)
    P, W = packed["P"], packed["W"]
    idx_w = np.zeros((P, Kp, 16), np.uint32)
    val_w = np.zeros((P, Kp, 16), np.uint32)
    for p in range(P):
        pk = np.ascontiguousarray(packed["packets"][p]).reshape(-1)
        oi = np.zeros(Kp * 16, np.uint32)
        ov = np.zeros(Kp * 16, np.uint32)
        L.orc_bscsr_partition_ex(pk, packed["packets"][p].shape[0], vec32, vec32.size, W, Kp, LFR, int(drift_free), oi, ov)
        idx_w[p] = oi.reshape(Kp, 16)
        val_w[p] = ov.reshape(Kp, 16)
    return idx_w, val_w


def read_result(idx_w, val_w, first_row, B):
    """host_spmv_bscsr.cpp:399-448 + evaluation_utils.hpp:40-62."""
    P, Kp, _ = idx_w.shape
    ri = np.zeros(P * Kp * 16, np.uint32)
    rv = np.zeros(P * Kp * 16, np.uint32)
    n = lib().orc_read_result(P, Kp, B, _c(idx_w, np.uint32).reshape(-1), _c(val_w, np.uint32).reshape(-1),
                              _c(first_row, np.uint32), ri, rv)
    return ri[:n].copy(), rv[:n].copy()


def bscsr_topk(row, col, val_f64, num_rows, vec_f32, P=32, W=20, Kp=8, LFR=4, drift_free=False):
    """Whole FPGA-mode pipeline on the CPU: quantise, pack, kernel, merge."""
    val32 = fx32_from_double(val_f64)
    vec32 = query_fx32_from_f32(vec_f32)
    packed = pack_bscsr(row, col, val32, num_rows, P, W)
    idx_w, val_w = bscsr_kernel(packed, vec32, Kp, LFR, drift_free)
    ri, rv = read_result(idx_w, val_w, packed["first_row"], packed["B"])
    return dict(idx=ri, val=rv, idx_words=idx_w, val_words=val_w, packed=packed, vec32=vec32, val32=val32)


def gold_topk_fx32(row, col, val32, vec32, k, sort=True):
    """gold_algorithms.hpp:188-246 with V = ap_ufixed<32,1> (host_spmv_bscsr.cpp:487-505)."""
    row, col = _c(row, np.uint32), _c(col, np.uint32)
    val32, vec32 = _c(val32, np.uint32), _c(vec32, np.uint32)
    idx = np.zeros(k, np.uint32)
    out = np.zeros(k, np.uint32)
    lib().orc_gold_topk_fx32(row, col, val32, row.size, vec32, k, idx, out)
    if sort:
        lib().orc_sort_tuples_u32(k, idx, out)
    return idx, out


# --------------------------------------------------------------------------
# the reference itself, compiled here (oracle/_ref) -- optional
# --------------------------------------------------------------------------

_ref_gold = None


def ref_gold():
    """oracle/_ref/libref_gold.so or None when it was never built."""
    global _ref_gold
    if _ref_gold is None and _REF_GOLD_PATH.exists():
        R = C.CDLL(str(_REF_GOLD_PATH))
        R.ref_gold_topk_f32.argtypes = [_u32p, _u32p, _f32p, C.c_uint64, _f32p, C.c_int, C.c_int, _u32p, _f32p]
        R.ref_sort_tuples_f32.argtypes = [C.c_uint32, _u32p, _f32p]
        R.ref_read_mtx.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                   C.POINTER(C.c_uint64)]
        R.ref_read_mtx.restype = C.c_int
        R.ref_read_mtx_fetch.argtypes = [_u32p, _u32p, _f32p]
        R.ref_coo2csr.argtypes = [_u32p, _u32p, _f32p, C.c_uint64, C.c_uint32, C.c_uint32, _u32p, _u32p, _f32p]
        R.ref_create_sample_vector_f32.argtypes = [_f32p, C.c_int, C.c_int]
        R.ref_coo_num_rows.argtypes = [_u32p, C.c_uint64]
        R.ref_coo_num_rows.restype = C.c_uint32
        R.ref_mean.argtypes = [_f32p, C.c_int, C.c_int]
        R.ref_mean.restype = C.c_float
        R.ref_st_dev.argtypes = [_f32p, C.c_int, C.c_int]
        R.ref_st_dev.restype = C.c_float
        _ref_gold = R
    return _ref_gold


def ref_gold_topk_f32(row, col, val, vec, k, sort=True):
    R = ref_gold()
    row, col, val, vec = _c(row, np.uint32), _c(col, np.uint32), _c(val, np.float32), _c(vec, np.float32)
    idx = np.zeros(k, np.uint32)
    out = np.zeros(k, np.float32)
    R.ref_gold_topk_f32(row, col, val, row.size, vec, k, 1 if sort else 0, idx, out)
    return idx, out


def ref_read_mtx(path, zero_indexed=False, sort=False):
    R = ref_gold()
    rows, cols, nnz = C.c_uint32(), C.c_uint32(), C.c_uint64()
    rc = R.ref_read_mtx(str(path).encode(), int(zero_indexed), int(sort), C.byref(rows), C.byref(cols),
                        C.byref(nnz))
    x = np.zeros(nnz.value, np.uint32)
    y = np.zeros(nnz.value, np.uint32)
    v = np.zeros(nnz.value, np.float32)
    R.ref_read_mtx_fetch(x, y, v)
    return rc, rows.value, cols.value, x, y, v


# --------------------------------------------------------------------------
# the reference's FPGA host + HLS kernel, compiled against oracle/shim and run in
# software (oracle/ref_fpga.cpp) -- optional, one library per knob combination
# --------------------------------------------------------------------------

_ref_fpga = {}


def ref_fpga_path(W=20, Kp=8, LFR=4, P=32):
    return _HERE / "_ref" / f"libref_fpga_w{W}_k{Kp}_l{LFR}_p{P}.so"


def ref_fpga(W=20, Kp=8, LFR=4, P=32):
    """ctypes handle of oracle/_ref/libref_fpga_w<W>_k<Kp>_l<LFR>_p<P>.so, or None when never built."""
    key = (W, Kp, LFR, P)
    if key not in _ref_fpga:
        path = ref_fpga_path(*key)
        if not path.exists():
            return None
        R = C.CDLL(str(path))
        ip = C.POINTER(C.c_int)
        R.ref_fpga_params.argtypes = [ip, ip, ip, ip, ip]
        R.ref_fx32_from_double.argtypes = [C.c_double]
        R.ref_fx32_from_double.restype = C.c_uint32
        R.ref_fxW_from_fx32.argtypes = [C.c_uint32]
        R.ref_fxW_from_fx32.restype = C.c_uint32
        R.ref_create_sample_vector_fx32.argtypes = [_u32p, C.c_int, C.c_int]
        R.ref_fpga_create.argtypes = [_u32p, _u32p, _u32p, C.c_uint64, C.c_uint32, C.c_uint32, _u32p]
        R.ref_fpga_create.restype = C.c_void_p
        R.ref_fpga_destroy.argtypes = [C.c_void_p]
        u32 = C.POINTER(C.c_uint32)
        R.ref_fpga_partition_info.argtypes = [C.c_void_p, C.c_int, u32, u32, u32, u32]
        R.ref_fpga_packets.argtypes = [C.c_void_p, C.c_int, _u64p]
        R.ref_fpga_query_blocks.argtypes = [C.c_void_p, _u64p]
        R.ref_fpga_run.argtypes = [C.c_void_p]
        R.ref_fpga_reset.argtypes = [C.c_void_p, _u32p]
        R.ref_fpga_result_words.argtypes = [C.c_void_p, C.c_int, _u32p, _u32p]
        R.ref_fpga_read_result.argtypes = [C.c_void_p, _u32p, _u32p, C.c_uint32]
        R.ref_fpga_read_result.restype = C.c_uint32
        R.ref_gold_topk_fx32.argtypes = [_u32p, _u32p, _u32p, C.c_uint64, _u32p, C.c_uint32, C.c_int, _u32p, _u32p]
        w, b, k, l, p = (C.c_int() for _ in range(5))
        R.ref_fpga_params(C.byref(w), C.byref(b), C.byref(k), C.byref(l), C.byref(p))
        assert (w.value, k.value, l.value, p.value) == key, "library knobs do not match its file name"
        R.packet_size = b.value
        _ref_fpga[key] = R
    return _ref_fpga[key]


class RefFpga:
    """The reference's `struct SpMV` (host_spmv_bscsr.cpp:79-485) driving its own HLS kernel in software."""

    def __init__(self, row, col, val32, num_rows, num_cols, vec32, W=20, Kp=8, LFR=4, P=32):
        self.R = ref_fpga(W, Kp, LFR, P)
        if self.R is None:
            raise FileNotFoundError(ref_fpga_path(W, Kp, LFR, P))
        self.W, self.Kp, self.LFR, self.P, self.B = W, Kp, LFR, P, self.R.packet_size
        row, col, val32, vec32 = _c(row, np.uint32), _c(col, np.uint32), _c(val32, np.uint32), _c(vec32, np.uint32)
        self.cols = int(num_cols)
        self.h = self.R.ref_fpga_create(row, col, val32, row.size, int(num_rows), self.cols, vec32)

    def close(self):
        if self.h:
            self.R.ref_fpga_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def partition_info(self):
        out = np.zeros((self.P, 4), np.uint32)   # first_row, last_row, nnz, num_blocks
        for p in range(self.P):
            v = [C.c_uint32() for _ in range(4)]
            self.R.ref_fpga_partition_info(self.h, p, *[C.byref(t) for t in v])
            out[p] = [t.value for t in v]
        return out

    def packets(self):
        info = self.partition_info()
        res = []
        for p in range(self.P):
            o = np.zeros((int(info[p, 3]), 8), np.uint64)
            self.R.ref_fpga_packets(self.h, p, o.reshape(-1))
            res.append(o)
        return res

    def query_blocks(self):
        o = np.zeros(((self.cols + self.B - 1) // self.B, 8), np.uint64)
        self.R.ref_fpga_query_blocks(self.h, o.reshape(-1))
        return o

    def run(self):
        self.R.ref_fpga_run(self.h)

    def reset(self, vec32):
        self.R.ref_fpga_reset(self.h, _c(vec32, np.uint32))

    def result_words(self):
        iw = np.zeros((self.P, self.Kp, 16), np.uint32)
        vw = np.zeros((self.P, self.Kp, 16), np.uint32)
        for p in range(self.P):
            a = np.zeros(self.Kp * 16, np.uint32)
            b = np.zeros(self.Kp * 16, np.uint32)
            self.R.ref_fpga_result_words(self.h, p, a, b)
            iw[p], vw[p] = a.reshape(self.Kp, 16), b.reshape(self.Kp, 16)
        return iw, vw

    def read_result(self):
        cap = self.P * self.Kp * 16
        ri, rv = np.zeros(cap, np.uint32), np.zeros(cap, np.uint32)
        n = self.R.ref_fpga_read_result(self.h, ri, rv, cap)
        return ri[:n].copy(), rv[:n].copy()
