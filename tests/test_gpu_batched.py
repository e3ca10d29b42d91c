"""GPU parity tests of the batched multi-query mode (BASELINE config 5: one matrix read serves 32 queries),
through the C ABI.

The batched kernel accumulates every row score sequentially in non-zero order with separate fp32 multiply
and add -- the arithmetic of the reference gold (gold_algorithms.hpp:203-213) -- so its scores must be
BIT-IDENTICAL to the oracle's sequential fp32 row sums; index sets may differ from a reference list only
where the k-th score is tied."""
import numpy as np
import pytest

from conftest import make_query
from test_gpu_csr import check_against_scores, exact_matrix

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cfg1(gen):
    x, y, v = gen.create_sparse_matrix(10000, 1024, 20, "gamma", seed=0)
    return x, y, v.astype(np.float32), gen.csr_from_coo(x, 10000)


def queries(cols, n, seed0=100):
    return np.stack([make_query(cols, seed0 + i) for i in range(n)])


def check_bit_exact(idx, val, cnt, yref, k):
    """scores bit-identical to the sequential fp32 sums; the list is the exact top-k under (score desc, index asc)
    except that rows tied with the k-th score may be exchanged for each other."""
    n = yref.size
    want = min(k, n)
    assert cnt == want
    idx, val = idx[:cnt], val[:cnt]
    assert np.array_equal(val.view(np.uint32), yref[idx].view(np.uint32)), "score differs from the sequential fp32 sum"
    order = np.lexsort((np.arange(n), -yref.astype(np.float64)))[:want]
    assert np.array_equal(val.view(np.uint32), yref[order].view(np.uint32))
    assert np.array_equal(idx, order.astype(np.uint32)), "total order (score desc, index asc) violated"


@pytest.mark.parametrize("batch", [2, 5, 31, 32, 33, 64, 100])
def test_batched_bit_exact_vs_sequential_fp32(cuda_required, tks, orc, cfg1, batch):
    x, y, v, ptr = cfg1
    Q = queries(1024, batch)
    with tks.SpMV(ptr, y, v, 10000, 1024, k=100, max_batch=128) as s:
        s.reset(Q)
        s()
        st = s.stats()
        assert st.launches_per_run == 5, "the amortised batched path must have run"
        assert st.batched_fallbacks == 0
        for q in range(batch):
            val, idx, cnt = s.read_result(q)
            check_bit_exact(idx, val, cnt, orc.spmv_f32(x, y, v, Q[q], 10000), 100)


def test_batched_matches_reference_gold(cuda_required, tks, orc, cfg1):
    """The reference's own streaming gold on the same inputs: identical score lists, bit for bit."""
    x, y, v, ptr = cfg1
    Q = queries(1024, 8, seed0=7)
    with tks.SpMV(ptr, y, v, 10000, 1024, k=100, max_batch=8) as s:
        s.reset(Q)
        s()
        for q in range(8):
            val, idx, cnt = s.read_result(q)
            gi, gv = orc.gold_topk_f32(x, y, v, Q[q], 100)
            assert np.array_equal(val.view(np.uint32), gv.view(np.uint32))
            assert set(idx.tolist()) == set(gi.tolist())


def test_batched_equals_one_pass_per_query(cuda_required, tks, orc, cfg1):
    x, y, v, ptr = cfg1
    Q = queries(1024, 40, seed0=900)
    with tks.SpMV(ptr, y, v, 10000, 1024, k=100, max_batch=64) as a, \
            tks.SpMV(ptr, y, v, 10000, 1024, k=100, max_batch=64, batch_mode=1) as b:
        a.reset(Q); a()
        b.reset(Q); b()
        assert b.stats().launches_per_run == 3 * 40
        for q in range(40):
            va, ia, ca = a.read_result(q)
            vb, ib, cb = b.read_result(q)
            assert ca == cb == 100
            np.testing.assert_allclose(va, vb, rtol=1e-5)       # different fp32 summation order
            assert len(set(ia.tolist()) ^ set(ib.tolist())) <= 2


@pytest.mark.parametrize("k", [1, 8, 128, 1024])
def test_batched_k_sweep(cuda_required, tks, orc, cfg1, k):
    x, y, v, ptr = cfg1
    Q = queries(1024, 33, seed0=k)
    with tks.SpMV(ptr, y, v, 10000, 1024, k=k, max_batch=64) as s:
        s.reset(Q)
        s()
        for q in (0, 15, 32):
            val, idx, cnt = s.read_result(q)
            check_bit_exact(idx, val, cnt, orc.spmv_f32(x, y, v, Q[q], 10000), k)


@pytest.mark.parametrize("tie_higher", [False, True])
@pytest.mark.parametrize("rows,k,max_deg,empty", [(5000, 100, 12, 0.0), (3000, 64, 3, 0.2), (50, 100, 5, 0.0),
                                                   (1, 8, 4, 0.0), (300, 7, 300, 0.0)])
def test_batched_exact_ties_and_edge_shapes(cuda_required, tks, orc, rows, k, max_deg, empty, tie_higher):
    """Massive ties, empty rows, fewer rows than k, a single row, rows longer than a staging batch: exact
    including the tie order, for every query of the batch."""
    rng = np.random.default_rng(rows * 17 + k)
    ptr, x, col, val, _ = exact_matrix(rows, 256, rng, max_deg, empty)
    Q = (2.0 ** rng.integers(-4, 1, (35, 256))).astype(np.float32)
    nonempty = np.diff(ptr.astype(np.int64)) > 0
    cand = np.nonzero(nonempty)[0]
    with tks.SpMV(ptr, col, val, rows, 256, k=k, tie_higher=tie_higher, chunk_nnz=256, max_batch=35) as s:
        s.reset(Q)
        s()
        assert s.stats().launches_per_run == 5
        for q in range(35):
            v, i, c = s.read_result(q)
            yref = orc.spmv_f32(x, col, val, Q[q], rows)
            want = min(k, cand.size)
            assert c == want
            sec = cand if not tie_higher else -cand
            order = cand[np.lexsort((sec, -yref[cand].astype(np.float64)))[:want]]
            assert np.array_equal(i[:c], order.astype(np.uint32))
            assert np.array_equal(v[:c].view(np.uint32), yref[order].view(np.uint32))


def test_batched_pool_overflow_falls_back_per_query(cuda_required, tks, orc, cfg1):
    """A pool far too small for the candidates: every query is re-run alone on the GPU and the result is
    still the exact top-k."""
    x, y, v, ptr = cfg1
    Q = queries(1024, 6, seed0=55)
    with tks.SpMV(ptr, y, v, 10000, 1024, k=100, max_batch=8, batch_pool_cap=16) as s:
        s.reset(Q)
        s()
        assert s.stats().batched_fallbacks >= 1
        for q in range(6):
            val, idx, cnt = s.read_result(q)
            check_against_scores(idx, val, cnt, orc.spmv_f32(x, y, v, Q[q], 10000), 100)
        # the async path resolves the overflow when the results are asked for
        s.reset(Q[::-1].copy())
        s.run_async(100)
        for q in range(6):
            val, idx, cnt = s.read_result(q)
            check_against_scores(idx, val, cnt, orc.spmv_f32(x, y, v, Q[5 - q], 10000), 100)


def test_batched_fma_variant_within_tolerance(cuda_required, tks, orc, cfg1):
    x, y, v, ptr = cfg1
    Q = queries(1024, 32, seed0=3)
    with tks.SpMV(ptr, y, v, 10000, 1024, k=100, max_batch=32, batch_fma=True) as s:
        s.reset(Q)
        s()
        for q in (0, 31):
            val, idx, cnt = s.read_result(q)
            check_against_scores(idx, val, cnt, orc.spmv_f32(x, y, v, Q[q], 10000), 100)


def test_batched_wide_matrix_uses_one_pass_per_query(cuda_required, tks, orc):
    """cols beyond what the 32-query table can hold in shared memory: the engine runs one pass per query."""
    rng = np.random.default_rng(4)
    rows, cols = 3000, 4000
    ptr, x, col, val, _ = exact_matrix(rows, cols, rng, 9)
    Q = queries(cols, 3)
    with tks.SpMV(ptr, col, val, rows, cols, k=20, max_batch=4, max_cols=cols) as s:
        s.reset(Q)
        s()
        assert s.stats().launches_per_run == 9
        for q in range(3):
            v, i, c = s.read_result(q)
            check_against_scores(i, v, c, orc.spmv_f32(x, col, val, Q[q], rows), 20)


def test_batched_large_synthetic_and_repeatability(cuda_required, tks, orc):
    """2M x 1024 gamma-20 generated in HBM (larger than L2), 64 queries, two runs: identical results, and
    bit-exact scores against the oracle for a few queries."""
    rows, cols = 2_000_000, 1024
    Q = queries(cols, 64, seed0=2000)
    with tks.SpMV(num_cols=cols, k=100, max_batch=64) as s:
        s.generate_synthetic(rows, cols, 20, "gamma", seed=11)
        ptr, idx, val = s.download_csr()
        x = np.repeat(np.arange(rows, dtype=np.uint32), np.diff(ptr.astype(np.int64)))
        s.reset(Q)
        s()
        first = [s.read_result(q) for q in range(64)]
        assert s.stats().batched_fallbacks == 0
        s()
        for q in range(64):
            v2, i2, c2 = s.read_result(q)
            assert np.array_equal(first[q][1], i2) and np.array_equal(first[q][0].view(np.uint32), v2.view(np.uint32))
        for q in (0, 31, 32, 63):
            v, i, c = first[q]
            check_bit_exact(i, v, c, orc.spmv_f32(x, idx, val, Q[q], rows), 100)


def test_batched_row_shards_merge_on_device(cuda_required, tks, orc, cfg1):
    """The multi-GPU batched path on one device: 2 row shards, keys regrouped by query as the all-gather
    would deliver them, one batched merge launch."""
    import torch
    x, y, v, ptr = cfg1
    sh = tks.sharding
    B, k = 33, 100
    Q = queries(1024, B, seed0=400)
    shards = sh.plan_row_shards_by_nnz(ptr, 2)
    engines, per_rank = [], []
    for r0, r1 in shards:
        p, i, vv = sh.slice_csr(ptr, y, v, r0, r1)
        e = tks.SpMV(p, i, vv, r1 - r0, 1024, k=k, row_offset=r0, max_batch=64)
        e.reset(Q)
        e.run_async(k)
        kp, _ = e.result_keys_device(0)
        kmax = 1024

        class _A:
            __cuda_array_interface__ = {"shape": (B, kmax), "typestr": "<i8", "data": (kp, False), "version": 2}
        per_rank.append(torch.as_tensor(_A(), device="cuda")[:, :k].clone())
        engines.append(e)
    torch.cuda.synchronize()
    gathered = torch.stack(per_rank, 0).permute(1, 0, 2).contiguous()      # [B][world][k]
    engines[0].merge_keys_batched_device(gathered.data_ptr(), 2 * k, B, k)
    for q in range(B):
        val, idx, cnt = engines[0].read_result(q)
        check_bit_exact(idx, val, cnt, orc.spmv_f32(x, y, v, Q[q], 10000), k)
    for e in engines:
        e.close()
