"""What scripts/sanitize.sh runs under compute-sanitizer: one small pass through every kernel family of the engine --
float (blocking and pipelined submits), 16-bit values, batched, the peer-window exchange between two shards of one
process (tks_group_*), and the fixed-point BS-CSR pipeline -- each checked against the oracle so that a sanitizer-clean
run is also a correct one."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
import oracle  # noqa: E402
from _pkg import pkg  # noqa: E402

tks = pkg()
gen = tks.create_matrices
rows, cols, k = 20000, 1024, 100
x, y, v = gen.create_sparse_matrix(rows, cols, 20, "gamma", seed=0)
v32 = v.astype(np.float32)
ptr = gen.csr_from_coo(x, rows)


def query(seed):
    r = np.random.default_rng(seed).random(cols)
    return (r / np.linalg.norm(r)).astype(np.float32)


def check(val, idx, q):
    gi, gv = oracle.gold_topk_f32(x, y, v32, q, k)
    np.testing.assert_allclose(val, gv, rtol=1e-5)
    assert len(set(idx.tolist()) ^ set(gi.tolist())) <= 2


# float: blocking verbs, then pipelined submits fed from the host
with tks.SpMV(ptr, y, v32, rows, cols, k=k) as s:
    for i in range(3):
        q = query(i)
        s.reset(q)
        s()
        val, idx, _ = s.read_result()
        check(val, idx, q)
    tickets = [s.submit_host(query(10 + i), k) for i in range(4)]
    for i, t in enumerate(tickets):
        val, idx, _ = s.fetch(t)
        check(val, idx, query(10 + i))
print("float ok")

# half-precision values, batched queries
with tks.SpMV(ptr, y, v32, rows, cols, k=k, half=True) as s:
    s.reset(query(3))
    s()
    s.read_result()
with tks.SpMV(ptr, y, v32, rows, cols, k=k, max_batch=40) as s:
    Q = np.stack([query(20 + i) for i in range(40)])
    s.reset(Q)
    s()
    val, idx, _ = s.read_result(39)
    check(val, idx, Q[39])
print("half / batched ok")

# two shards of one process exchanging candidates through the peer windows inside the select kernels
L = tks.capi.lib()
cfg = tks.capi.default_config(mode=tks.capi.MODE_FLOAT_CSR)
devs = np.zeros(2, np.int32)
g = C.c_void_p()
assert L.tks_group_create(C.byref(cfg), devs.ctypes.data_as(C.c_void_p), 2, C.byref(g)) == 0
p64 = ptr.astype(np.uint64)
assert L.tks_group_upload_csr(g, rows, cols, y.size, p64.ctypes.data_as(C.c_void_p), 64, y.ctypes.data_as(C.c_void_p), v32.ctypes.data_as(C.c_void_p)) == 0
idx, val, cnt = np.zeros(1024, np.uint32), np.zeros(1024, np.float32), C.c_uint32()
for i in range(3):
    q = query(30 + i)
    assert L.tks_group_set_query(g, q.ctypes.data_as(C.c_void_p)) == 0
    assert L.tks_group_run(g, k, None, None) == 0, L.tks_group_last_error(g)
    assert L.tks_group_read_result(g, 1, idx.ctypes.data_as(C.c_void_p), val.ctypes.data_as(C.c_void_p), C.byref(cnt)) == 0
    check(val[:k], idx[:k], q)
L.tks_group_destroy(g)
print("group exchange ok")

# fixed-point BS-CSR engine (FPGA semantics), bit-exact against the oracle
q = query(40)
with tks.SpMVFixed(x, y, oracle.fx32_from_double(v), rows, cols, vec32=oracle.query_fx32_from_f32(q), k=k) as f:
    f()
    fv, fi = f.read_result()
o = oracle.bscsr_topk(x, y, v, rows, q)
assert np.array_equal(fi, o["idx"][:k]) and np.array_equal(fv, o["val"][:k])
print("fixed ok")
