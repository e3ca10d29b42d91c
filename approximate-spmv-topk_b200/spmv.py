"""Python mirror of the reference's `struct SpMV` functor over the C ABI.

Same four verbs as src/fpga/src/host_spmv_bscsr.cpp:79-485 and src/gpu/host_spmv_topk_csr_gpu.cu:44-263:
constructor (upload), __call__ (run, returns kernel nanoseconds), read_result, reset.  All arrays are
HOST numpy arrays; every call goes through libtopkspmv.so -- there is no NumPy compute path here.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import _ptr, check


class _Base:
    handle = None

    def _create(self, cfg):
        L = capi.lib()
        h = C.c_void_p()
        rc = L.tks_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise capi.TksError(rc, (L.tks_last_error(None) or b"").decode())
        self.handle = h
        self.cfg = cfg

    def close(self):
        if self.handle is not None:
            capi.lib().tks_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def stats(self) -> capi.TksStats:
        st = capi.TksStats()
        check(capi.lib().tks_get_stats(self.handle, C.byref(st)), self.handle)
        return st

    def set_profile_kernels(self, on=True):
        check(capi.lib().tks_set_profile_kernels(self.handle, int(bool(on))), self.handle)

    def run_timed(self, k=None):
        """operator() with both timings: (kernel_ms, total_ms)."""
        k = self.k if k is None else k
        km, tm = C.c_float(), C.c_float()
        check(capi.lib().tks_run(self.handle, k, C.byref(km), C.byref(tm)), self.handle)
        self.k = k
        return km.value, tm.value

    def __call__(self, debug=0):
        km, _ = self.run_timed()
        return km * 1e6

    def run_async(self, k=None, stream=0):
        k = self.k if k is None else k
        check(capi.lib().tks_run_async(self.handle, k, C.c_void_p(stream)), self.handle)
        self.k = k


class SpMV(_Base):
    """Exact fp32 CSR engine (host_spmv_topk_csr_gpu.cu:95 signature: ptr, idx, val, rows, cols, nnz, vec, k)."""

    def __init__(self, ptr=None, idx=None, val=None, num_rows=0, num_cols=0, num_nnz=None, vec=None, k=100,
                 device=0, tie_higher=False, max_batch=1, max_cols=None, chunk_nnz=0, row_offset=0,
                 profile_kernels=False, batch_mode=0, batch_pool_cap=0, batch_fma=False, half=False, bf16=False):
        # half: the reference's use_half_precision_gpu (host_spmv_topk_csr_gpu.cu:95): values and query rounded to
        # IEEE half; this engine still accumulates in fp32
        cfg = capi.default_config(mode=capi.MODE_FLOAT_CSR, device=device, max_batch=max_batch,
                                  tie_break=capi.TIE_HIGHER_INDEX if tie_higher else capi.TIE_LOWER_INDEX,
                                  max_cols=max(1024, int(max_cols or num_cols or 1024)), chunk_nnz=chunk_nnz,
                                  profile_kernels=int(profile_kernels), batch_mode=int(batch_mode),
                                  batch_pool_cap=int(batch_pool_cap), batch_fma=int(batch_fma),
                                  value_type=capi.VALUE_BF16 if bf16 else (capi.VALUE_FP16 if half else capi.VALUE_FP32))
        self._create(cfg)
        self.k = k
        self.num_rows, self.num_cols = int(num_rows), int(num_cols)
        if ptr is not None:
            self.upload(ptr, idx, val, num_rows, num_cols, row_offset)
            if vec is not None:
                self.reset(vec)

    def upload(self, ptr, idx, val, num_rows, num_cols, row_offset=0):
        ptr = np.ascontiguousarray(ptr)
        if ptr.dtype not in (np.uint32, np.uint64):
            ptr = ptr.astype(np.uint64)
        idx = np.ascontiguousarray(idx, np.uint32)
        val = np.ascontiguousarray(val, np.float32)
        bits = 32 if ptr.dtype == np.uint32 else 64
        check(capi.lib().tks_upload_csr(self.handle, int(num_rows), int(num_cols), idx.size, _ptr(ptr), bits,
                                        _ptr(idx), _ptr(val), int(row_offset)), self.handle)
        self.num_rows, self.num_cols = int(num_rows), int(num_cols)

    def generate_synthetic(self, rows, cols, avg_degree, dist="gamma", seed=0, row_offset=0):
        d = {"uniform": 0, "gamma": 1}[dist]
        check(capi.lib().tks_generate_synthetic(self.handle, int(rows), int(cols), int(avg_degree), d, int(seed),
                                                int(row_offset)), self.handle)
        self.num_rows, self.num_cols = int(rows), int(cols)

    def download_csr(self):
        st = self.stats()
        ptr = np.zeros(st.rows + 1, np.uint64)
        idx = np.zeros(st.nnz, np.uint32)
        val = np.zeros(st.nnz, np.float32)
        check(capi.lib().tks_download_csr(self.handle, _ptr(ptr), _ptr(idx), _ptr(val)), self.handle)
        return ptr, idx, val

    def download_csr_rows(self, row_begin, row_end):
        """(ptr uint64[n+1] rebased to 0, idx, val) of the resident rows [row_begin, row_end)."""
        n = int(row_end) - int(row_begin)
        ptr = np.zeros(n + 1, np.uint64)
        check(capi.lib().tks_download_csr_rows(self.handle, int(row_begin), int(row_end), _ptr(ptr), None, None), self.handle)
        nnz = int(ptr[-1] - ptr[0])
        idx = np.zeros(max(nnz, 1), np.uint32)
        val = np.zeros(max(nnz, 1), np.float32)
        check(capi.lib().tks_download_csr_rows(self.handle, int(row_begin), int(row_end), None, _ptr(idx), _ptr(val)), self.handle)
        return ptr - ptr[0], idx[:nnz], val[:nnz]

    def reset(self, vec, debug=0):
        vec = np.ascontiguousarray(vec, np.float32)
        batch = 1 if vec.ndim == 1 else vec.shape[0]
        assert vec.shape[-1] == self.num_cols, "query length must equal num_cols"
        if batch == 1:
            # one query: through a staging array whose ctypes pointer is made once (building `.ctypes` for a fresh array
            # costs more host time than copying 4 KB)
            st = self.__dict__.get("_q_stage")
            if st is None or st[0].size != self.num_cols:
                buf = np.empty(self.num_cols, np.float32)
                st = self._q_stage = (buf, _ptr(buf))
            st[0][:] = vec.reshape(-1)
            qp = st[1]
        else:
            qp = _ptr(vec)
        check(capi.lib().tks_set_query(self.handle, qp, batch), self.handle)
        self.batch = batch
        return 0

    def reset_device(self, dptr, batch=1, stream=0):
        check(capi.lib().tks_set_query_device(self.handle, C.c_void_p(dptr), batch, C.c_void_p(stream)), self.handle)
        self.batch = batch

    def read_result(self, query=0):
        """Returns (values float32[k], indices uint32[k], count)."""
        out = self.__dict__.get("_out_stage")
        if out is None or out[0].size < self.k:
            idx, val, cnt = np.zeros(max(self.k, 1024), np.uint32), np.zeros(max(self.k, 1024), np.float32), C.c_uint32()
            out = self._out_stage = (idx, val, cnt, _ptr(idx), _ptr(val), C.byref(cnt))   # pointers made once
        check(capi.lib().tks_read_result(self.handle, query, out[3], out[4], out[5]), self.handle)
        return out[1][:self.k].copy(), out[0][:self.k].copy(), out[2].value

    def result_keys_device(self, query=0):
        p, n = C.c_void_p(), C.c_uint32()
        check(capi.lib().tks_result_keys_device(self.handle, query, C.byref(p), C.byref(n)), self.handle)
        return p.value, n.value

    # ---- candidate exchange over peer memory (several GPUs of one box) ----
    def peer_init(self, world, rank) -> bytes:
        buf = (C.c_uint8 * capi.IPC_HANDLE_BYTES)()
        check(capi.lib().tks_peer_init(self.handle, world, rank, C.cast(buf, C.c_void_p)), self.handle)
        return bytes(buf)

    def peer_connect(self, handles):
        blob = b"".join(handles)
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        check(capi.lib().tks_peer_connect(self.handle, C.cast(buf, C.c_void_p)), self.handle)

    def run_exchange_async(self, k=None, stream=0):
        k = self.k if k is None else k
        check(capi.lib().tks_run_exchange_async(self.handle, k, C.c_void_p(stream)), self.handle)
        self.k = k

    def peer_exchange_async(self, k=None, stream=0):
        k = self.k if k is None else k
        check(capi.lib().tks_peer_exchange_async(self.handle, k, C.c_void_p(stream)), self.handle)

    # ---- pipelined submits: consecutive queries overlap (tks_submit) ----
    def submit(self, dptr, k=None, stream=0, exchange=False, query_ready=False):
        """Enqueue one query (device pointer to num_cols fp32 values, read in place).  At most four are in flight."""
        k = self.k if k is None else k
        flags = (capi.SUBMIT_EXCHANGE if exchange else 0) | (capi.SUBMIT_QUERY_READY if query_ready else 0)
        check(capi.lib().tks_submit(self.handle, C.c_void_p(dptr), k, flags, C.c_void_p(stream)), self.handle)
        self.k = k
        self.batch = 1

    def submit_host(self, vec, k=None, exchange=False):
        """The pipeline fed from host memory: enqueue `vec` (num_cols fp32 values), return a ticket for fetch().
        The throughput form of reset(vec) + operator(): up to four queries are in flight."""
        k = self.k if k is None else k
        st = self.__dict__.get("_sh_stage")
        if st is None or st[0].size != self.num_cols:
            buf = np.empty(self.num_cols, np.float32)
            st = self._sh_stage = (buf, _ptr(buf), C.c_uint64())
        st[0][:] = np.asarray(vec, np.float32).reshape(-1)
        check(capi.lib().tks_submit_host(self.handle, st[1], k, capi.SUBMIT_EXCHANGE if exchange else 0, C.byref(st[2])),
              self.handle)
        self.k = k
        self.batch = 1
        return st[2].value

    def fetch(self, ticket):
        """read_result() of the query submit_host() gave `ticket` for: (values float32[k], indices uint32[k], count)."""
        out = self.__dict__.get("_fetch_stage")
        if out is None:
            idx, val, cnt = np.zeros(1024, np.uint32), np.zeros(1024, np.float32), C.c_uint32()
            out = self._fetch_stage = (idx, val, cnt, _ptr(idx), _ptr(val), C.byref(cnt))
        check(capi.lib().tks_fetch(self.handle, int(ticket), out[3], out[4], out[5]), self.handle)
        return out[1][:self.k].copy(), out[0][:self.k].copy(), out[2].value

    def pipeline_wait(self, stream=0):
        """Make `stream` wait (on the device) for the result of the last submitted query."""
        check(capi.lib().tks_pipeline_wait(self.handle, C.c_void_p(stream)), self.handle)

    def pipeline_stamps(self, n=1024):
        """%globaltimer nanoseconds of the last n submits, oldest first: uint64[n, 8] with columns sample begin / end,
        main begin / end, select resident / begin / end (see topkspmv.h)."""
        buf = np.zeros((n, 8), np.uint64)
        cnt = C.c_uint32()
        check(capi.lib().tks_pipeline_stamps(self.handle, _ptr(buf), n, C.byref(cnt)), self.handle)
        return buf[:cnt.value].copy()

    def merge_keys_device(self, dptr, n_keys, k=None, query=0, stream=0):
        k = self.k if k is None else k
        check(capi.lib().tks_merge_keys_device(self.handle, query, C.c_void_p(dptr), n_keys, k, C.c_void_p(stream)),
              self.handle)


    def merge_keys_batched_device(self, dptr, keys_per_query, batch, k=None, stream=0):
        k = self.k if k is None else k
        check(capi.lib().tks_merge_keys_batched_device(self.handle, C.c_void_p(dptr), keys_per_query, batch, k,
                                                       C.c_void_p(stream)), self.handle)


class SpMVFixed(_Base):
    """FPGA-semantics engine (host_spmv_bscsr.cpp:104 signature: x, y, val, rows, cols, nnz, vec).

    x, y: row-sorted COO; val32 / vec32: raw ap_ufixed<32,1> words (real_type_inout)."""

    def __init__(self, x, y, val32, num_rows, num_cols, vec32=None, k=100, fixed_width=20, partitions=32, local_k=8,
                 limited_finished_rows=4, device=0, profile_kernels=False, drift_free=False, device_pack=False):
        # device_pack: partitioning, packets and device tables are built on the GPU from the COO arrays
        # (tks_upload_coo_fixed) instead of tks_pack_bscsr on the host + tks_upload_bscsr; same resident state
        cfg = capi.default_config(mode=capi.MODE_FIXED_BSCSR, device=device, fixed_width=fixed_width,
                                  partitions=partitions, local_k=local_k,
                                  limited_finished_rows=limited_finished_rows, tie_break=capi.TIE_HIGHER_INDEX,
                                  profile_kernels=int(profile_kernels), fixed_drift_free=int(drift_free))
        self._create(cfg)
        self.k = k
        self.num_rows, self.num_cols = int(num_rows), int(num_cols)
        self.B = capi.bscsr_packet_size(fixed_width)
        if device_pack:
            self.upload_coo(x, y, val32, num_rows, num_cols)
        else:
            packets, ppp, first, npp = capi.pack_bscsr(x, y, val32, num_rows, partitions, fixed_width)
            self.upload_packets(packets, ppp, first, npp)
        if vec32 is not None:
            self.reset(vec32)

    def upload_coo(self, x, y, val32, num_rows, num_cols):
        x = np.ascontiguousarray(x, np.uint32)
        y = np.ascontiguousarray(y, np.uint32)
        val32 = np.ascontiguousarray(val32, np.uint32)
        check(capi.lib().tks_upload_coo_fixed(self.handle, _ptr(x), _ptr(y), _ptr(val32), x.size, int(num_rows),
                                              int(num_cols)), self.handle)
        self.partitions = int(self.cfg.partitions)
        # first row of every partition = row of the first non-zero of its row range (host:136-145)
        rpp = (int(num_rows) + self.partitions - 1) // self.partitions
        starts = np.searchsorted(x, np.arange(self.partitions, dtype=np.int64) * rpp, side="left")
        self.first_row = x[starts].astype(np.uint32)

    def first_row_array(self):
        """First row of every partition (host_spmv_bscsr.cpp:145), local to this engine's matrix."""
        return np.asarray(self.first_row, np.uint32)

    def state_digest(self):
        d = np.zeros(16, np.uint64)
        check(capi.lib().tks_bscsr_state_digest(self.handle, _ptr(d), 16), self.handle)
        return d

    def upload_packets(self, packets, ppp, first_row, npp):
        P = len(ppp)
        self._packets = packets   # keep alive during the call
        offs = np.concatenate([[0], np.cumsum(ppp)]).astype(np.uint64)
        base = packets.ctypes.data
        ptrs = (C.c_void_p * P)(*[C.c_void_p(base + int(offs[p]) * 64) for p in range(P)])
        ppp = np.ascontiguousarray(ppp, np.uint64)
        first_row = np.ascontiguousarray(first_row, np.uint32)
        npp = np.ascontiguousarray(npp, np.uint64)
        check(capi.lib().tks_upload_bscsr(self.handle, self.num_cols, P, _ptr(ppp), C.cast(ptrs, C.c_void_p),
                                          _ptr(first_row), _ptr(npp)), self.handle)
        self.first_row = first_row
        self.partitions = P

    def reset(self, vec32, debug=0):
        vec32 = np.ascontiguousarray(vec32, np.uint32)
        assert vec32.size == self.num_cols
        st = self.__dict__.get("_q_stage")          # staging array whose ctypes pointer is made once
        if st is None:
            buf = np.empty(self.num_cols, np.uint32)
            st = self._q_stage = (buf, _ptr(buf))
        st[0][:] = vec32.reshape(-1)
        check(capi.lib().tks_set_query(self.handle, st[1], 1), self.handle)
        return 0

    def reset_device(self, dptr, stream=0):
        check(capi.lib().tks_set_query_device(self.handle, C.c_void_p(dptr), 1, C.c_void_p(stream)), self.handle)

    def submit(self, dptr, k=None, stream=0, query_ready=False):
        """Pipelined reset_device + run_async for a query (raw words) already in HBM: its transform and sample run on
        a second stream beside the previous query's stream and replay kernels; read_result() returns the last one."""
        k = self.k if k is None else k
        check(capi.lib().tks_submit(self.handle, C.c_void_p(dptr), k, capi.SUBMIT_QUERY_READY if query_ready else 0,
                                    C.c_void_p(stream)), self.handle)
        self.k = k

    def submit_host(self, vec32, k=None):
        """Pipelined reset(vec) + operator(): the query's transform, copy and sample overlap the previous query's stream
        and replay kernels; returns a ticket for fetch().  Two queries are kept."""
        k = self.k if k is None else k
        st = self.__dict__.get("_sh_stage")
        if st is None:
            buf = np.empty(self.num_cols, np.uint32)
            st = self._sh_stage = (buf, _ptr(buf), C.c_uint64())
        st[0][:] = np.asarray(vec32, np.uint32).reshape(-1)
        check(capi.lib().tks_submit_host(self.handle, st[1], k, 0, C.byref(st[2])), self.handle)
        self.k = k
        return st[2].value

    def fetch(self, ticket):
        """read_result() of the query submit_host() gave `ticket` for: (raw values uint32[n], indices uint32[n]), n <= k."""
        out = self.__dict__.get("_fetch_stage")
        if out is None:
            idx, val, cnt = np.zeros(1024, np.uint32), np.zeros(1024, np.uint32), C.c_uint32()
            out = self._fetch_stage = (idx, val, cnt, _ptr(idx), _ptr(val), C.byref(cnt))
        check(capi.lib().tks_fetch(self.handle, int(ticket), out[3], out[4], out[5]), self.handle)
        n = out[2].value
        return out[1][:n].copy(), out[0][:n].copy()

    def read_result(self):
        """Returns (raw values uint32[count], indices uint32[count]) -- may be shorter than k."""
        out = self.__dict__.get("_out_stage")
        if out is None or out[0].size < self.k:
            idx, val, cnt = np.zeros(max(self.k, 1024), np.uint32), np.zeros(max(self.k, 1024), np.uint32), C.c_uint32()
            out = self._out_stage = (idx, val, cnt, _ptr(idx), _ptr(val), C.byref(cnt))   # pointers made once
        check(capi.lib().tks_read_result(self.handle, 0, out[3], out[4], out[5]), self.handle)
        n = out[2].value
        return out[1][:n].copy(), out[0][:n].copy()

    def partition_words_device(self):
        """(device pointer, word count) of the last async run's result words: index words, then value words."""
        p, n = C.c_void_p(), C.c_uint32()
        check(capi.lib().tks_partition_words_device(self.handle, C.byref(p), C.byref(n)), self.handle)
        return p.value, n.value

    def read_partition_results(self):
        Kp = self.cfg.local_k
        idx_w = np.zeros((self.partitions, Kp, 16), np.uint32)
        val_w = np.zeros((self.partitions, Kp, 16), np.uint32)
        check(capi.lib().tks_read_partition_results(self.handle, _ptr(idx_w), _ptr(val_w)), self.handle)
        return idx_w, val_w
