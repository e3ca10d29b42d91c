"""tks_group_*: several GPUs (or several shards of one GPU) driven by ONE process through the C ABI -- the single-process
counterpart of the torch.distributed plumbing (SURVEY 8b `num_gpus / device_ids`, 8e).  Every member must hold the
same global list, bit for bit; against the one-device engine the bar is the float bar of the north star (scores within
1e-5 relative, index sets equal except near-ties at the k-th place): a shard's rows sit at other lane offsets than in
the unsharded stream, so their fp32 sums may associate differently."""
import ctypes as C

import numpy as np
import pytest

from conftest import make_query

pytestmark = pytest.mark.gpu


def _same(val, idx, cnt, wv, wi, wc, k):
    assert cnt == wc
    np.testing.assert_allclose(np.sort(val[:k])[::-1], np.sort(wv)[::-1], rtol=1e-5, atol=1e-7)
    assert len(set(idx[:k].tolist()) ^ set(wi.tolist())) <= (2 if k <= 100 else 6)


def _devices(n):
    import torch
    have = torch.cuda.device_count()
    return [i % have for i in range(n)]


@pytest.mark.parametrize("n,k", [(2, 100), (3, 100), (4, 300), (8, 100), (2, 1024)])
def test_group_equals_single_engine(cuda_required, tks, orc, gen, n, k):
    rows, cols = 120_000, 1024
    x, y, v = gen.create_sparse_matrix(rows, cols, 20, "gamma", seed=31)
    v = v.astype(np.float32)
    ptr = gen.csr_from_coo(x, rows)
    queries = [make_query(cols, 700 + i) for i in range(6)]
    want = []
    with tks.SpMV(ptr, y, v, rows, cols, k=k) as s:
        for q in queries:
            s.reset(q)
            s()
            want.append(s.read_result())
    L = tks.capi.lib()
    cfg = tks.capi.default_config(mode=tks.capi.MODE_FLOAT_CSR)
    devs = np.asarray(_devices(n), np.int32)
    g = C.c_void_p()
    rc = L.tks_group_create(C.byref(cfg), devs.ctypes.data_as(C.c_void_p), n, C.byref(g))
    assert rc == 0, L.tks_group_last_error(None)
    try:
        assert L.tks_group_size(g) == n
        p64 = ptr.astype(np.uint64)
        assert L.tks_group_upload_csr(g, rows, cols, y.size, p64.ctypes.data_as(C.c_void_p), 64, y.ctypes.data_as(C.c_void_p),
                                      v.ctypes.data_as(C.c_void_p)) == 0, L.tks_group_last_error(g)
        idx, val, cnt = np.zeros(1024, np.uint32), np.zeros(1024, np.float32), C.c_uint32()
        for rep in range(2):                                   # the windows' parity slots are reused
            for q, (wv, wi, wc) in zip(queries, want):
                assert L.tks_group_set_query(g, q.ctypes.data_as(C.c_void_p)) == 0
                km, tm = C.c_float(), C.c_float()
                assert L.tks_group_run(g, k, C.byref(km), C.byref(tm)) == 0, L.tks_group_last_error(g)
                first = None
                for member in range(n):                        # every member holds the same global result
                    assert L.tks_group_read_result(g, member, idx.ctypes.data_as(C.c_void_p), val.ctypes.data_as(C.c_void_p), C.byref(cnt)) == 0
                    if first is None:
                        first = (val[:k].copy(), idx[:k].copy(), cnt.value)
                        _same(val, idx, cnt.value, wv, wi, wc, k)
                    else:
                        assert cnt.value == first[2] and np.array_equal(idx[:k], first[1]) and np.array_equal(val[:k], first[0]), f"member {member}"
        # the pipelined, host-fed form over the group
        tickets = []
        got = {}
        t = C.c_uint64()
        for i, q in enumerate(queries):
            assert L.tks_group_submit_host(g, q.ctypes.data_as(C.c_void_p), k, C.byref(t)) == 0, L.tks_group_last_error(g)
            tickets.append(t.value)
            if i >= 2:
                assert L.tks_group_fetch(g, tickets[i - 2], idx.ctypes.data_as(C.c_void_p), val.ctypes.data_as(C.c_void_p), C.byref(cnt)) == 0, L.tks_group_last_error(g)
                got[i - 2] = (val[:k].copy(), idx[:k].copy(), cnt.value)
        for i in range(len(queries) - 2, len(queries)):
            assert L.tks_group_fetch(g, tickets[i], idx.ctypes.data_as(C.c_void_p), val.ctypes.data_as(C.c_void_p), C.byref(cnt)) == 0
            got[i] = (val[:k].copy(), idx[:k].copy(), cnt.value)
        for i, (wv, wi, wc) in enumerate(want):
            _same(got[i][0], got[i][1], got[i][2], wv, wi, wc, k)
    finally:
        L.tks_group_destroy(g)
    # and against the oracle
    yref = orc.spmv_f32(x, y, v, queries[0], rows)
    np.testing.assert_allclose(want[0][0], yref[want[0][1]], rtol=1e-5)


def test_group_rejects_bad_arguments(tks):
    L = tks.capi.lib()
    cfg = tks.capi.default_config(mode=tks.capi.MODE_FIXED_BSCSR)
    devs = np.zeros(2, np.int32)
    g = C.c_void_p()
    assert L.tks_group_create(C.byref(cfg), devs.ctypes.data_as(C.c_void_p), 2, C.byref(g)) == tks.capi.TKS_EINVAL
    cfg = tks.capi.default_config(mode=tks.capi.MODE_FLOAT_CSR)
    assert L.tks_group_create(C.byref(cfg), devs.ctypes.data_as(C.c_void_p), 9, C.byref(g)) == tks.capi.TKS_EINVAL
