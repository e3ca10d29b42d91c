"""Pipelined submits (tks_submit): consecutive queries overlap on the device -- the sample of query i+1 beside the
main kernel of query i, the main kernels chained without waiting, the select on a third stream -- and every result
must still be the one the un-pipelined path (tks_run_async, same kernels in stream order) gives, bit for bit, and the
one the oracle gives within the float tolerance."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

from conftest import make_query

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
RTOL = 1e-5


class _DevView:
    def __init__(self, ptr, shape, typestr="<i8"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 2}


@pytest.fixture(scope="module")
def big(gen):
    """400k rows: long enough (8 M non-zeros) for the kernels of consecutive queries to really overlap."""
    rows = 400_000
    x, y, v = gen.create_sparse_matrix(rows, 1024, 20, "gamma", seed=11)
    return x, y, v.astype(np.float32), gen.csr_from_coo(x, rows), rows


def _plain_results(tks, ptr, y, v, rows, queries, k, **kw):
    out = []
    with tks.SpMV(ptr, y, v, rows, 1024, k=k, **kw) as s:
        for q in queries:
            s.reset(q)
            s.run_async(k)
            val, idx, cnt = s.read_result()
            out.append((val, idx, cnt))
    return out


@pytest.mark.parametrize("k", [1, 100, 129, 500, 1024])
def test_every_query_of_a_deep_pipeline_equals_the_unpipelined_run(cuda_required, tks, orc, big, k):
    import torch
    x, y, v, ptr, rows = big
    n = 24
    queries = np.stack([make_query(1024, 900 + i) for i in range(n)])
    want = _plain_results(tks, ptr, y, v, rows, queries, k)
    dq = torch.from_numpy(queries).cuda()
    stream = torch.cuda.Stream()
    with tks.SpMV(ptr, y, v, rows, 1024, k=k) as s, torch.cuda.stream(stream):
        kp, _ = s.result_keys_device(0)
        keys = torch.as_tensor(_DevView(kp, (1024,)), device="cuda")
        snaps = torch.zeros((n, k), dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        for i in range(n):
            s.submit(dq[i].data_ptr(), k, stream.cuda_stream, query_ready=True)
            # snapshot of query i's keys once its select has run, ordered before the next main kernel of the stream
            s.pipeline_wait(stream.cuda_stream)
            snaps[i].copy_(keys[:k], non_blocking=True)
        val, idx, cnt = s.read_result()
        torch.cuda.synchronize()
        got = snaps.cpu().numpy().view(np.uint64)
        stamps = s.pipeline_stamps(n)
    assert cnt == k and np.array_equal(idx, want[-1][1]) and np.array_equal(val, want[-1][0])
    assert stamps.shape == (n, 8) and np.all(np.diff(stamps[:, 6].astype(np.int64)) > 0), "select kernels did not finish in order"
    st = stamps.astype(np.int64)
    assert np.all(st[:, 1] >= st[:, 0]) and np.all(st[:, 3] >= st[:, 2]) and np.all(st[:, 5] >= st[:, 3]) and np.all(st[:, 6] >= st[:, 5])
    assert np.all(st[:, 2] >= st[:, 1]), "a main kernel started streaming before its threshold was published"
    for i in range(n):
        wv, wi, wc = want[i]
        score = (got[i] >> np.uint64(32)).astype(np.uint32)
        row = (~got[i]).astype(np.uint32)           # TKS_TIE_LOWER_INDEX keys hold ~row
        assert np.array_equal(row[:wc], wi[:wc]), f"query {i}: indices differ from the un-pipelined run"
        b = np.where(score & 0x80000000, score & 0x7FFFFFFF, ~score).astype(np.uint32)
        assert np.array_equal(b[:wc], wv[:wc].view(np.uint32)), f"query {i}: scores differ from the un-pipelined run"
    # and against the oracle for a few of them
    for i in (0, n // 2, n - 1):
        yref = orc.spmv_f32(x, y, v, queries[i], rows)
        np.testing.assert_allclose(want[i][0], yref[want[i][1]], rtol=RTOL, atol=1e-7)
        kth = np.sort(yref)[::-1][k - 1]
        must = np.nonzero(yref > kth * (1 + RTOL) + 1e-7)[0]
        assert set(must.tolist()) <= set(want[i][1].tolist())


def test_free_running_pipeline_last_result_and_mixing_with_plain_runs(cuda_required, tks, big):
    """No waits between submits (the benchmark's loop); then a blocking run, then submits again."""
    import torch
    x, y, v, ptr, rows = big
    k, n = 100, 40
    queries = np.stack([make_query(1024, 1300 + i) for i in range(n)])
    want = _plain_results(tks, ptr, y, v, rows, queries, k)
    dq = torch.from_numpy(queries).cuda()
    torch.cuda.synchronize()
    with tks.SpMV(ptr, y, v, rows, 1024, k=k) as s:
        for rounds in range(2):
            for i in range(n):
                s.submit(dq[i].data_ptr(), k, 0, query_ready=True)
            val, idx, cnt = s.read_result()
            assert cnt == k and np.array_equal(idx, want[n - 1][1]) and np.array_equal(val, want[n - 1][0])
            assert s.stats().last_candidates > 0
            s.reset(queries[3])
            s()
            val, idx, cnt = s.read_result()
            assert np.array_equal(idx, want[3][1]) and np.array_equal(val, want[3][0])
            # a query produced on the caller's stream right before the submit (no QUERY_READY)
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                buf = torch.empty(1024, dtype=torch.float32, device="cuda")
                buf.copy_(torch.from_numpy(queries[7]).pin_memory(), non_blocking=True)
                s.submit(buf.data_ptr(), k, stream.cuda_stream)
            val, idx, cnt = s.read_result()
            assert np.array_equal(idx, want[7][1]) and np.array_equal(val, want[7][0])


@pytest.mark.parametrize("mode", ["half", "bf16", "empty_rows_tie_higher"])
def test_pipeline_in_the_other_modes(cuda_required, tks, gen, mode):
    import torch
    rows = 60_000
    x, y, v = gen.create_sparse_matrix(rows, 1024, 20, "gamma", seed=21)
    v = v.astype(np.float32)
    kw = {}
    if mode == "empty_rows_tie_higher":
        keep = (x % 7) != 3                          # every seventh row empty
        x, y, v = x[keep], y[keep], v[keep]
        kw = dict(tie_higher=True)
    elif mode == "half":
        kw = dict(half=True)
    else:
        kw = dict(bf16=True)
    ptr = gen.csr_from_coo(x, rows)
    k, n = 100, 9
    queries = np.stack([make_query(1024, 50 + i) for i in range(n)])
    want = _plain_results(tks, ptr, y, v, rows, queries, k, **kw)
    dq = torch.from_numpy(queries).cuda()
    torch.cuda.synchronize()
    with tks.SpMV(ptr, y, v, rows, 1024, k=k, **kw) as s:
        for i in range(n):
            s.submit(dq[i].data_ptr(), k, 0, query_ready=True)
            if i % 4 == 3 or i == n - 1:
                val, idx, cnt = s.read_result()
                assert cnt == want[i][2] and np.array_equal(idx, want[i][1]) and np.array_equal(val, want[i][0])


@pytest.mark.parametrize("k", [100, 1024])
def test_host_fed_pipeline_equals_the_blocking_verbs(cuda_required, tks, big, k):
    """tks_submit_host / tks_fetch: host query in, host result out, up to four queries in flight; every result equals
    reset(vec) + operator() + read_result() for the same query, bit for bit."""
    x, y, v, ptr, rows = big
    n = 30
    queries = np.stack([make_query(1024, 2100 + i) for i in range(n)])
    want = _plain_results(tks, ptr, y, v, rows, queries, k)
    with tks.SpMV(ptr, y, v, rows, 1024, k=k) as s:
        tickets, got = [], {}
        for i in range(n):
            tickets.append(s.submit_host(queries[i], k))
            if i >= 3:
                got[i - 3] = s.fetch(tickets[i - 3])
        for i in range(n - 3, n):
            got[i] = s.fetch(tickets[i])
        for i in range(n):
            val, idx, cnt = got[i]
            assert cnt == want[i][2] and np.array_equal(idx, want[i][1]) and np.array_equal(val, want[i][0]), f"query {i}"
        # a result that has been pushed out of the four slots is refused, not silently replaced
        t_old = s.submit_host(queries[0], k)
        for i in range(1, 5):
            s.submit_host(queries[i], k)
        with pytest.raises(tks.capi.TksError, match="gone"):
            s.fetch(t_old)
        with pytest.raises(tks.capi.TksError, match="tks_fetch"):
            s.read_result()
        # the blocking verbs still work afterwards
        s.reset(queries[5])
        s()
        val, idx, cnt = s.read_result()
        assert np.array_equal(idx, want[5][1]) and np.array_equal(val, want[5][0])


def test_submit_rejects_bad_calls(cuda_required, tks, gen):
    x, y, v = gen.create_sparse_matrix(2000, 1024, 20, "gamma", seed=1)
    ptr = gen.csr_from_coo(x, 2000)
    with tks.SpMV(num_cols=1024) as s:
        with pytest.raises(tks.capi.TksError):
            s.submit(1 << 20, 100)                   # no matrix
    with tks.SpMV(ptr, y, v.astype(np.float32), 2000, 1024) as s:
        with pytest.raises(tks.capi.TksError):
            s.submit(1 << 20, 0)                     # k out of range
        with pytest.raises(tks.capi.TksError):
            s.submit(1 << 20, 100, exchange=True)    # no peers connected


# ---- several GPUs: pipelined submits with the exchange fused into the select kernel ------------------------------

def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q, k, n):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    from _pkg import pkg
    tks = pkg()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    gen, sh = tks.create_matrices, tks.sharding
    rows, cols = 300_000, 1024
    x, y, v = gen.create_sparse_matrix(rows, cols, 20, "gamma", seed=5)
    v = v.astype(np.float32)
    ptr = gen.csr_from_coo(x, rows)
    r0, r1 = sh.plan_row_shards_by_nnz(ptr, world)[rank]
    p, idx, val = sh.slice_csr(ptr, y, v, r0, r1)
    Q = np.stack([make_query(cols, 300 + i) for i in range(n)])
    dq = torch.from_numpy(Q).cuda()
    stream = torch.cuda.Stream()
    out = []
    with torch.cuda.stream(stream):
        eng = tks.SpMV(p, idx, val, r1 - r0, cols, k=k, device=rank, row_offset=r0)
        sharded = tks.ShardedSpMV(eng, k, batch=1, exchange="auto")
        torch.cuda.synchronize()
        dist.barrier()
        for i in range(n):
            sharded.submit(dq[i].data_ptr(), stream.cuda_stream)
            if i % 5 == 4 or i == n - 1:             # free-running in between
                a, b, c = eng.read_result()
                out.append((i, a.copy(), b.copy(), c))
        # the host-fed form of the same steps: every step's global result is fetched, three steps behind
        tickets = [None] * n
        for i in range(n):
            tickets[i] = sharded.submit_host(Q[i])
            if i >= 3:
                a, b, c = sharded.fetch(tickets[i - 3])
                out.append((i - 3, a.copy(), b.copy(), c))
        for i in range(n - 3, n):
            a, b, c = sharded.fetch(tickets[i])
            out.append((i, a.copy(), b.copy(), c))
    q.put((rank, out, sharded.exchange_mode))
    dist.barrier()
    eng.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("k", [100, 300])
def test_pipelined_exchange_over_ranks_equals_global_topk(cuda_required, tks, orc, gen, k):
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 8)
    if world > 2 and world * k > 2048:
        world = 2
    n = 17
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, k, n)) for r in range(world)]
    for p in procs:
        p.start()
    from conftest import collect_from_workers
    results = {r: (out, mode) for r, out, mode in collect_from_workers(q, procs, world, 600)}
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    rows, cols = 300_000, 1024
    x, y, v = gen.create_sparse_matrix(rows, cols, 20, "gamma", seed=5)
    v = v.astype(np.float32)
    if world > 1:
        assert {results[r][1] for r in results} == {"peer"}
    yrefs = {}
    for j, (i, _, _, _) in enumerate(results[0][0]):
        if i not in yrefs:
            yrefs[i] = orc.spmv_f32(x, y, v, make_query(cols, 300 + i), rows)
        yref = yrefs[i]
        order = np.lexsort((np.arange(rows), -yref.astype(np.float64)))[:k]
        for rank in range(world):
            _, val, idx, cnt = results[rank][0][j]
            assert cnt == k
            assert len(set(idx.tolist()) ^ set(order.tolist())) <= (0 if k <= 100 else 4)
            np.testing.assert_allclose(val, yref[idx], rtol=RTOL)
            assert np.array_equal(idx, results[0][0][j][2]) and np.array_equal(val, results[0][0][j][1]), "ranks disagree"
