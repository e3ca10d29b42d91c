"""ctypes binding of libtopkspmv.so -- one Python function per entry point of include/topkspmv.h.

This is what a maintainer of the (C++/Python) reference would add to call the engine from
test_cpu.py-style drivers; see INTEGRATION.md.  Loading fails loudly when the library has not been
built: there is no Python/NumPy fallback for any compute call.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "lib" / "libtopkspmv.so"

TKS_OK, TKS_EINVAL, TKS_ECUDA, TKS_ESTATE, TKS_ENOMEM, TKS_EIO = 0, -1, -2, -3, -4, -5
MODE_FLOAT_CSR, MODE_FIXED_BSCSR = 0, 1
TIE_LOWER_INDEX, TIE_HIGHER_INDEX = 0, 1
VALUE_FP32, VALUE_FP16, VALUE_BF16 = 0, 1, 2
IPC_HANDLE_BYTES = 128
SUBMIT_EXCHANGE, SUBMIT_QUERY_READY = 1, 2


class TksConfig(C.Structure):
    _fields_ = [("mode", C.c_int32), ("fixed_width", C.c_int32), ("partitions", C.c_int32),
                ("local_k", C.c_int32), ("limited_finished_rows", C.c_int32), ("max_cols", C.c_int32),
                ("tie_break", C.c_int32), ("device", C.c_int32), ("max_batch", C.c_int32),
                ("chunk_nnz", C.c_int32), ("profile_kernels", C.c_int32), ("batch_mode", C.c_int32),
                ("batch_pool_cap", C.c_int32), ("batch_fma", C.c_int32), ("fixed_drift_free", C.c_int32),
                ("value_type", C.c_int32)]


class TksStats(C.Structure):
    _fields_ = [("rows", C.c_uint64), ("cols", C.c_uint64), ("nnz", C.c_uint64), ("packets", C.c_uint64),
                ("algorithmic_bytes", C.c_uint64), ("device_bytes", C.c_uint64),
                ("last_kernel_ms", C.c_float), ("last_total_ms", C.c_float),
                ("last_candidates", C.c_uint32), ("launches_per_run", C.c_uint32),
                ("last_main_kernel_ms", C.c_float), ("batched_fallbacks", C.c_uint32),
                ("logged_candidates", C.c_uint32), ("work_unit_nnz", C.c_uint32), ("work_units", C.c_uint32),
                ("reserved", C.c_uint32 * 3)]


class TksError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libtopkspmv error {code}: {msg}")
        self.code = code


# every symbol include/topkspmv.h declares (tests/test_capi_symbols.py checks the .so exports them all)
SYMBOLS = [
    "tks_version", "tks_default_config", "tks_create", "tks_destroy", "tks_last_error",
    "tks_upload_csr", "tks_upload_csr_device", "tks_upload_bscsr", "tks_generate_synthetic",
    "tks_upload_coo_fixed", "tks_upload_coo_fixed_device", "tks_bscsr_state_digest",
    "tks_download_csr", "tks_download_csr_rows", "tks_set_query", "tks_set_query_device", "tks_run", "tks_run_async",
    "tks_read_result", "tks_read_partition_results", "tks_partition_words_device", "tks_result_keys_device", "tks_merge_keys_device", "tks_merge_keys_batched_device",
    "tks_peer_init", "tks_peer_connect", "tks_run_exchange_async", "tks_peer_exchange_async",
    "tks_group_create", "tks_group_destroy", "tks_group_last_error", "tks_group_size", "tks_group_member",
    "tks_group_upload_csr", "tks_group_generate_synthetic", "tks_group_set_query", "tks_group_run", "tks_group_read_result",
    "tks_group_submit_host", "tks_group_fetch",
    "tks_submit", "tks_submit_host", "tks_fetch", "tks_pipeline_wait", "tks_pipeline_stamps",
    "tks_set_profile_kernels", "tks_get_stats", "tks_bscsr_packet_size", "tks_fixed32_from_double", "tks_fixedW_from_fixed32",
    "tks_pack_bscsr", "tks_merge_partition_words", "tks_read_mtx", "tks_coo2csr",
    "tks_cache_write_csr", "tks_cache_read_csr", "tks_cache_source_tag", "tks_cache_write_csr_tagged", "tks_cache_read_csr_tagged", "tks_cache_write_bscsr", "tks_cache_read_bscsr",
]

_lib = None


def lib() -> C.CDLL:
    """Load libtopkspmv.so (built in-tree by build.py).  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise FileNotFoundError(
            f"{LIB_PATH} not found: build it with `python approximate-spmv-topk_b200/build.py` "
            "(the engine has no CPU fallback)")
    L = C.CDLL(str(LIB_PATH))
    vp, u32p, u64p, f32p = C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_float)
    L.tks_version.restype = C.c_int
    L.tks_default_config.argtypes = [C.POINTER(TksConfig)]
    L.tks_create.argtypes = [C.POINTER(TksConfig), C.POINTER(vp)]
    L.tks_destroy.argtypes = [vp]
    L.tks_destroy.restype = None
    L.tks_last_error.argtypes = [vp]
    L.tks_last_error.restype = C.c_char_p
    L.tks_host_last_error.restype = C.c_char_p
    L.tks_upload_csr.argtypes = [vp, C.c_uint64, C.c_uint32, C.c_uint64, vp, C.c_int, vp, vp, C.c_uint64]
    L.tks_upload_csr_device.argtypes = [vp, C.c_uint64, C.c_uint32, C.c_uint64, vp, C.c_int, vp, vp, C.c_uint64]
    L.tks_upload_bscsr.argtypes = [vp, C.c_uint32, C.c_uint32, vp, vp, vp, vp]
    L.tks_upload_coo_fixed.argtypes = [vp, vp, vp, vp, C.c_uint64, C.c_uint32, C.c_uint32]
    L.tks_upload_coo_fixed_device.argtypes = [vp, vp, vp, vp, C.c_uint64, C.c_uint32, C.c_uint32]
    L.tks_bscsr_state_digest.argtypes = [vp, vp, C.c_uint32]
    L.tks_generate_synthetic.argtypes = [vp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.c_uint64, C.c_uint64]
    L.tks_download_csr.argtypes = [vp, vp, vp, vp]
    L.tks_download_csr_rows.argtypes = [vp, C.c_uint64, C.c_uint64, vp, vp, vp]
    L.tks_set_query.argtypes = [vp, vp, C.c_uint32]
    L.tks_set_query_device.argtypes = [vp, vp, C.c_uint32, vp]
    L.tks_run.argtypes = [vp, C.c_uint32, f32p, f32p]
    L.tks_run_async.argtypes = [vp, C.c_uint32, vp]
    L.tks_read_result.argtypes = [vp, C.c_uint32, vp, vp, u32p]
    L.tks_read_partition_results.argtypes = [vp, vp, vp]
    L.tks_partition_words_device.argtypes = [vp, C.POINTER(vp), u32p]
    L.tks_result_keys_device.argtypes = [vp, C.c_uint32, C.POINTER(vp), u32p]
    L.tks_merge_keys_device.argtypes = [vp, C.c_uint32, vp, C.c_uint32, C.c_uint32, vp]
    L.tks_merge_keys_batched_device.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, vp]
    L.tks_peer_init.argtypes = [vp, C.c_uint32, C.c_uint32, vp]
    L.tks_peer_connect.argtypes = [vp, vp]
    L.tks_run_exchange_async.argtypes = [vp, C.c_uint32, vp]
    L.tks_peer_exchange_async.argtypes = [vp, C.c_uint32, vp]
    L.tks_group_create.argtypes = [C.POINTER(TksConfig), vp, C.c_uint32, C.POINTER(vp)]
    L.tks_group_destroy.argtypes = [vp]
    L.tks_group_destroy.restype = None
    L.tks_group_last_error.argtypes = [vp]
    L.tks_group_last_error.restype = C.c_char_p
    L.tks_group_size.argtypes = [vp]
    L.tks_group_size.restype = C.c_uint32
    L.tks_group_member.argtypes = [vp, C.c_uint32]
    L.tks_group_member.restype = vp
    L.tks_group_upload_csr.argtypes = [vp, C.c_uint64, C.c_uint32, C.c_uint64, vp, C.c_int, vp, vp]
    L.tks_group_generate_synthetic.argtypes = [vp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.c_uint64]
    L.tks_group_set_query.argtypes = [vp, vp]
    L.tks_group_run.argtypes = [vp, C.c_uint32, f32p, f32p]
    L.tks_group_read_result.argtypes = [vp, C.c_uint32, vp, vp, u32p]
    L.tks_group_submit_host.argtypes = [vp, vp, C.c_uint32, u64p]
    L.tks_group_fetch.argtypes = [vp, C.c_uint64, vp, vp, u32p]
    L.tks_submit.argtypes = [vp, vp, C.c_uint32, C.c_uint32, vp]
    L.tks_submit_host.argtypes = [vp, vp, C.c_uint32, C.c_uint32, u64p]
    L.tks_fetch.argtypes = [vp, C.c_uint64, vp, vp, u32p]
    L.tks_pipeline_wait.argtypes = [vp, vp]
    L.tks_pipeline_stamps.argtypes = [vp, vp, C.c_uint32, u32p]
    L.tks_set_profile_kernels.argtypes = [vp, C.c_int]
    L.tks_get_stats.argtypes = [vp, C.POINTER(TksStats)]
    L.tks_bscsr_packet_size.argtypes = [C.c_int]
    L.tks_fixed32_from_double.argtypes = [C.c_double]
    L.tks_fixed32_from_double.restype = C.c_uint32
    L.tks_fixedW_from_fixed32.argtypes = [C.c_uint32, C.c_int]
    L.tks_fixedW_from_fixed32.restype = C.c_uint32
    L.tks_pack_bscsr.argtypes = [vp, vp, vp, C.c_uint64, C.c_uint32, C.c_int, C.c_int, vp, vp, vp, vp]
    L.tks_merge_partition_words.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, vp, vp, vp, C.c_int, C.c_uint32, vp, vp, u32p]
    L.tks_read_mtx.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, u32p, u32p, u64p, C.c_uint64, vp, vp, vp]
    L.tks_coo2csr.argtypes = [vp, vp, vp, C.c_uint64, C.c_uint32, C.c_uint32, vp, vp, vp]
    _lib = L
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def check(rc, handle=None):
    if rc != 0:
        L = lib()
        msg = L.tks_last_error(handle) if handle else L.tks_host_last_error()
        raise TksError(rc, (msg or b"").decode(errors="replace"))


def default_config(**overrides) -> TksConfig:
    cfg = TksConfig()
    check(lib().tks_default_config(C.byref(cfg)))
    for k, v in overrides.items():
        if not hasattr(cfg, k):
            raise AttributeError(f"tks_config has no field {k}")
        setattr(cfg, k, v)
    return cfg


# ---- host-side helpers (no GPU needed) -------------------------------------------------------

def bscsr_packet_size(W: int) -> int:
    return int(lib().tks_bscsr_packet_size(W))


def fixed32_from_double(a) -> np.ndarray:
    a = np.asarray(a, np.float64)
    f = lib().tks_fixed32_from_double
    return np.fromiter((f(float(v)) for v in a.ravel()), np.uint32, a.size).reshape(a.shape)


def fixed32_from_double_np(a) -> np.ndarray:
    """Vectorised (T) value cast of utils.hpp:401 for arrays: double -> raw ap_ufixed<32,1,AP_TRN_ZERO>.
    Same rule as tks_fixed32_from_double (tests compare them); host-side data preparation only."""
    a = np.asarray(a, np.float64)
    s = np.floor(np.where(a > 0, a, 0.0) * 2147483648.0)
    return np.mod(s, 4294967296.0).astype(np.uint64).astype(np.uint32)


def pack_bscsr(row, col, val32, num_rows, partitions=32, fixed_width=20):
    """tks_pack_bscsr: returns (packets uint64[total,8], packets_per_part, first_row, nnz_per_part)."""
    row = np.ascontiguousarray(row, np.uint32)
    col = np.ascontiguousarray(col, np.uint32)
    val32 = np.ascontiguousarray(val32, np.uint32)
    ppp = np.zeros(partitions, np.uint64)
    first = np.zeros(partitions, np.uint32)
    npp = np.zeros(partitions, np.uint64)
    L = lib()
    check(L.tks_pack_bscsr(_ptr(row), _ptr(col), _ptr(val32), row.size, num_rows, partitions, fixed_width,
                           _ptr(ppp), _ptr(first), _ptr(npp), None))
    packets = np.zeros((int(ppp.sum()), 8), np.uint64)
    check(L.tks_pack_bscsr(_ptr(row), _ptr(col), _ptr(val32), row.size, num_rows, partitions, fixed_width,
                           _ptr(ppp), _ptr(first), _ptr(npp), _ptr(packets)))
    return packets, ppp, first, npp


def merge_partition_words(idx_words, val_words, first_row, packet_size, k, tie_break=TIE_HIGHER_INDEX):
    """tks_merge_partition_words: idx_words/val_words uint32[P, Kp, 16] -> (values uint32[n], indices uint32[n]), n <= k."""
    iw = np.ascontiguousarray(idx_words, np.uint32)
    vw = np.ascontiguousarray(val_words, np.uint32)
    fr = np.ascontiguousarray(first_row, np.uint32)
    P, Kp = iw.shape[0], iw.shape[1]
    assert iw.shape == vw.shape == (P, Kp, 16) and fr.size == P
    idx = np.zeros(k, np.uint32)
    val = np.zeros(k, np.uint32)
    cnt = C.c_uint32()
    check(lib().tks_merge_partition_words(P, Kp, int(packet_size), _ptr(iw), _ptr(vw), _ptr(fr), int(tie_break), k,
                                          _ptr(idx), _ptr(val), C.byref(cnt)))
    n = min(k, cnt.value)
    return val[:n], idx[:n]


def read_mtx(path, zero_indexed=False, sort_tuples=False, ignore_values=False):
    """tks_read_mtx: returns (rows, cols, x uint32, y uint32, val float64)."""
    L = lib()
    rows, cols, nnz = C.c_uint32(), C.c_uint32(), C.c_uint64()
    p = str(path).encode()
    check(L.tks_read_mtx(p, int(zero_indexed), int(sort_tuples), int(ignore_values), C.byref(rows), C.byref(cols),
                         C.byref(nnz), 0, None, None, None))
    n = nnz.value
    x = np.zeros(n, np.uint32)
    y = np.zeros(n, np.uint32)
    v = np.zeros(n, np.float64)
    check(L.tks_read_mtx(p, int(zero_indexed), int(sort_tuples), int(ignore_values), C.byref(rows), C.byref(cols),
                         C.byref(nnz), n, _ptr(x), _ptr(y), _ptr(v)))
    return rows.value, cols.value, x, y, v


def coo2csr(x, y, val, rows, cols):
    x = np.ascontiguousarray(x, np.uint32)
    y = np.ascontiguousarray(y, np.uint32)
    val = np.ascontiguousarray(val, np.float32)
    ptr = np.zeros(rows + 1, np.uint32)
    idx = np.zeros(x.size, np.uint32)
    out = np.zeros(x.size, np.float32)
    check(lib().tks_coo2csr(_ptr(x), _ptr(y), _ptr(val), x.size, rows, cols, _ptr(ptr), _ptr(idx), _ptr(out)))
    return ptr, idx, out


# ---- binary matrix cache --------------------------------------------------------------

def cache_write_csr(path, ptr, idx, val, cols):
    ptr = np.ascontiguousarray(ptr, np.uint64)
    idx = np.ascontiguousarray(idx, np.uint32)
    val = np.ascontiguousarray(val, np.float32)
    L = lib()
    L.tks_cache_write_csr.argtypes = [C.c_char_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    check(L.tks_cache_write_csr(str(path).encode(), ptr.size - 1, int(cols), idx.size, _ptr(ptr), _ptr(idx), _ptr(val)))


def cache_read_csr(path):
    """Returns (rows, cols, ptr uint64, idx uint32, val float32)."""
    L = lib()
    L.tks_cache_read_csr.argtypes = [C.c_char_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64),
                                     C.c_void_p, C.c_void_p, C.c_void_p]
    rows, cols, nnz = C.c_uint64(), C.c_uint32(), C.c_uint64()
    p = str(path).encode()
    check(L.tks_cache_read_csr(p, C.byref(rows), C.byref(cols), C.byref(nnz), None, None, None))
    ptr = np.zeros(rows.value + 1, np.uint64)
    idx = np.zeros(nnz.value, np.uint32)
    val = np.zeros(nnz.value, np.float32)
    check(L.tks_cache_read_csr(p, C.byref(rows), C.byref(cols), C.byref(nnz), _ptr(ptr), _ptr(idx), _ptr(val)))
    return rows.value, cols.value, ptr, idx, val


def cache_write_bscsr(path, rows, cols, fixed_width, packets, ppp, first_row, npp):
    packets = np.ascontiguousarray(packets, np.uint64)
    ppp = np.ascontiguousarray(ppp, np.uint64)
    first_row = np.ascontiguousarray(first_row, np.uint32)
    npp = np.ascontiguousarray(npp, np.uint64)
    L = lib()
    L.tks_cache_write_bscsr.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p]
    check(L.tks_cache_write_bscsr(str(path).encode(), int(rows), int(cols), int(fixed_width), ppp.size, _ptr(ppp),
                                  _ptr(first_row), _ptr(npp), _ptr(packets)))


def cache_read_bscsr(path):
    """Returns (rows, cols, fixed_width, packets uint64[total,8], packets_per_part, first_row, nnz_per_part)."""
    L = lib()
    L.tks_cache_read_bscsr.argtypes = [C.c_char_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_int),
                                       C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p]
    rows, cols, W, P, total = C.c_uint32(), C.c_uint32(), C.c_int(), C.c_uint32(), C.c_uint64()
    p = str(path).encode()
    check(L.tks_cache_read_bscsr(p, C.byref(rows), C.byref(cols), C.byref(W), C.byref(P), C.byref(total), None, None,
                                 None, None))
    ppp = np.zeros(P.value, np.uint64)
    first = np.zeros(P.value, np.uint32)
    npp = np.zeros(P.value, np.uint64)
    packets = np.zeros((total.value, 8), np.uint64)
    check(L.tks_cache_read_bscsr(p, C.byref(rows), C.byref(cols), C.byref(W), C.byref(P), C.byref(total), _ptr(ppp),
                                 _ptr(first), _ptr(npp), _ptr(packets)))
    return rows.value, cols.value, W.value, packets, ppp, first, npp
