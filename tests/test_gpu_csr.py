"""GPU parity tests of the fused fp32 CSR Top-K SpMV engine, through the C ABI (ctypes).

Bar (BASELINE.json north_star): scores within 1e-5 relative of the CPU reference, index sets identical
except where scores tie within that tolerance; with exactly-representable data the result must be
bit-exact including the tie-break order."""
import numpy as np
import pytest

from conftest import make_query

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def check_against_scores(idx, val, cnt, y, k, tie="lower", exact=False):
    """y: reference score of every row (fp32, sequential order = the reference gold's arithmetic)."""
    n = y.size
    want = min(k, n)
    assert cnt == want
    idx, val = idx[:cnt], val[:cnt]
    assert len(set(idx.tolist())) == cnt, "duplicate rows in the result"
    assert np.all(idx < n)
    assert np.all(np.diff(val) <= 0), "scores not sorted descending"
    if exact:
        order = np.lexsort((np.arange(n) if tie == "lower" else -np.arange(n), -y.astype(np.float64)))[:want]
        assert np.array_equal(idx, order.astype(np.uint32))
        assert np.array_equal(val.view(np.uint32), y[order].view(np.uint32))
        return
    np.testing.assert_allclose(val, y[idx], rtol=RTOL, atol=1e-7)
    ys = np.sort(y)[::-1]
    kth = ys[want - 1]
    must = np.nonzero(y > kth * (1 + RTOL) + 1e-7)[0]
    assert set(must.tolist()) <= set(idx.tolist()), "a row clearly above the k-th score is missing"
    assert np.all(y[idx] >= kth * (1 - RTOL) - 1e-7), "a row clearly below the k-th score was returned"


def run_engine(tks, ptr, col, val, rows, cols, vec, k, **kw):
    with tks.SpMV(ptr, col, val, rows, cols, vec=vec, k=k, **kw) as s:
        s()
        v, i, c = s.read_result()
    return i, v, c


@pytest.fixture(scope="module")
def cfg1(gen):
    """BASELINE config 1: matrix_10000_1024_20_gamma, k=100."""
    x, y, v = gen.create_sparse_matrix(10000, 1024, 20, "gamma", seed=0)
    return x, y, v.astype(np.float32), gen.csr_from_coo(x, 10000)


def test_cfg1_matches_gold_and_float64(cuda_required, tks, orc, cfg1):
    x, y, v, ptr = cfg1
    with tks.SpMV(ptr, y, v, 10000, 1024, k=100) as s:
        for qs in range(1, 6):
            vec = make_query(1024, qs)
            s.reset(vec)
            s()
            val, idx, cnt = s.read_result()
            yref = orc.spmv_f32(x, y, v, vec, 10000)
            check_against_scores(idx, val, cnt, yref, 100)
            gi, gv = orc.gold_topk_f32(x, y, v, vec, 100)          # the reference's own algorithm
            np.testing.assert_allclose(val, gv, rtol=RTOL)
            fi, fv = orc.f64_topk(ptr, y, v, vec, 100)             # sparse_dot_topn stand-in
            np.testing.assert_allclose(val, fv, rtol=RTOL)
            assert len(set(idx.tolist()) ^ set(fi.tolist())) <= 2   # only boundary near-ties may differ


@pytest.mark.parametrize("k", [1, 8, 20, 100, 128, 129, 384, 500, 1024])
def test_k_sweep(cuda_required, tks, orc, cfg1, k):
    x, y, v, ptr = cfg1
    vec = make_query(1024, 77)
    idx, val, cnt = run_engine(tks, ptr, y, v, 10000, 1024, vec, k)
    check_against_scores(idx, val, cnt, orc.spmv_f32(x, y, v, vec, 10000), k)


@pytest.mark.parametrize("chunk_nnz", [128, 256, 1024, 4096, 65536])
def test_chunk_sizes(cuda_required, tks, orc, cfg1, chunk_nnz):
    x, y, v, ptr = cfg1
    vec = make_query(1024, 5)
    idx, val, cnt = run_engine(tks, ptr, y, v, 10000, 1024, vec, 100, chunk_nnz=chunk_nnz)
    check_against_scores(idx, val, cnt, orc.spmv_f32(x, y, v, vec, 10000), 100)


def exact_matrix(rows, cols, rng, max_deg=12, empty_frac=0.0):
    """Values and query are small powers of two: every summation order gives the same fp32 result."""
    deg = rng.integers(1, max_deg + 1, rows)
    if empty_frac > 0:
        deg[rng.random(rows) < empty_frac] = 0
    ptr = np.zeros(rows + 1, np.uint64)
    np.cumsum(deg, out=ptr[1:])
    nnz = int(ptr[-1])
    x = np.repeat(np.arange(rows, dtype=np.uint32), deg)
    col = rng.integers(0, cols, nnz).astype(np.uint32)
    val = (2.0 ** rng.integers(-3, 1, nnz)).astype(np.float32)
    vec = (2.0 ** rng.integers(-4, 1, cols)).astype(np.float32)
    return ptr, x, col, val, vec


@pytest.mark.parametrize("tie_higher", [False, True])
@pytest.mark.parametrize("rows,k,max_deg,empty", [(5000, 100, 12, 0.0), (5000, 100, 3, 0.0), (3000, 64, 8, 0.2),
                                                   (50, 100, 5, 0.0), (1, 8, 4, 0.0), (300, 7, 300, 0.0)])
def test_exact_ties_and_edge_shapes(cuda_required, tks, orc, rows, k, max_deg, empty, tie_higher):
    """Massive ties (few distinct scores), rows of length 1..3 inside one lane, empty rows, fewer rows
    than k, a single row, rows longer than a warp iteration: bit-exact incl. tie order."""
    rng = np.random.default_rng(rows * 31 + k)
    ptr, x, col, val, vec = exact_matrix(rows, 256, rng, max_deg, empty)
    yref = orc.spmv_f32(x, col, val, vec, rows)
    nonempty = np.diff(ptr.astype(np.int64)) > 0
    with tks.SpMV(ptr, col, val, rows, 256, vec=vec, k=k, tie_higher=tie_higher, chunk_nnz=256) as s:
        s()
        v, i, c = s.read_result()
    # empty rows are never candidates (the reference's COO gold never sees them either)
    cand = np.nonzero(nonempty)[0]
    want = min(k, cand.size)
    assert c == want
    sec = cand if not tie_higher else -cand
    order = cand[np.lexsort((sec, -yref[cand].astype(np.float64)))[:want]]
    assert np.array_equal(i[:c], order.astype(np.uint32))
    assert np.array_equal(v[:c].view(np.uint32), yref[order].view(np.uint32))
    assert np.all(i[c:] == 0) and np.all(v[c:] == 0)


def test_negative_and_zero_scores_exact(cuda_required, tks, orc):
    rng = np.random.default_rng(3)
    ptr, x, col, val, vec = exact_matrix(4000, 128, rng, 6)
    val = val * rng.choice(np.array([-1.0, 1.0], np.float32), val.size)
    yref = orc.spmv_f32(x, col, val, vec, 4000)
    idx, v, cnt = run_engine(tks, ptr, col, val, 4000, 128, vec, 1000)
    check_against_scores(idx, v, cnt, yref, 1000, exact=True)


def test_long_rows_and_wide_matrix(cuda_required, tks, orc):
    """Rows far longer than a chunk, cols up to the 14-bit limit."""
    rng = np.random.default_rng(8)
    rows, cols = 64, 16383
    deg = rng.integers(1, 9000, rows)
    ptr = np.zeros(rows + 1, np.uint64)
    np.cumsum(deg, out=ptr[1:])
    nnz = int(ptr[-1])
    x = np.repeat(np.arange(rows, dtype=np.uint32), deg)
    col = rng.integers(0, cols, nnz).astype(np.uint32)
    val = rng.random(nnz).astype(np.float32)
    vec = make_query(cols, 1)
    idx, v, cnt = run_engine(tks, ptr, col, val, rows, cols, vec, 10, max_cols=cols)
    y64 = np.zeros(rows)
    np.add.at(y64, x, val.astype(np.float64) * vec[col].astype(np.float64))
    order = np.argsort(-y64)[:10]
    assert np.array_equal(np.sort(idx[:cnt]), np.sort(order.astype(np.uint32)))
    np.testing.assert_allclose(v[:cnt], y64[idx[:cnt]], rtol=2e-5)


def test_row_offset_and_repeatability(cuda_required, tks, orc, cfg1):
    x, y, v, ptr = cfg1
    vec = make_query(1024, 21)
    with tks.SpMV(ptr, y, v, 10000, 1024, vec=vec, k=50, row_offset=123456) as s:
        s()
        v1, i1, _ = s.read_result()
        for _ in range(5):
            s()
            v2, i2, _ = s.read_result()
            assert np.array_equal(i1, i2) and np.array_equal(v1.view(np.uint32), v2.view(np.uint32))
    idx0, val0, _ = run_engine(tks, ptr, y, v, 10000, 1024, vec, 50)
    assert np.array_equal(i1, idx0 + 123456)
    assert np.array_equal(v1.view(np.uint32), val0.view(np.uint32))


def test_errors(cuda_required, tks):
    ptr = np.array([0, 2, 4], np.uint64)
    col = np.array([0, 1, 2, 5000], np.uint32)
    val = np.ones(4, np.float32)
    with pytest.raises(tks.capi.TksError, match="column index"):
        tks.SpMV(ptr, col, val, 2, 1024)
    good_col = np.array([0, 1, 2, 3], np.uint32)
    # row_ptr must run from 0 to nnz: stray non-zeros in front of ptr[0] or behind ptr[rows] are rejected, not scored
    for bad_ptr in ([1, 2, 4], [0, 2, 3], [0, 3, 2]):
        with pytest.raises(tks.capi.TksError, match="row_ptr"):
            tks.SpMV(np.array(bad_ptr, np.uint64), good_col, val, 2, 1024)
    with tks.SpMV(ptr, good_col, val, 2, 1024, k=100) as s:
        with pytest.raises(tks.capi.TksError, match="no query"):
            s()
        s.reset(np.ones(1024, np.float32))
        with pytest.raises(tks.capi.TksError):
            s.run_timed(k=5000)


def test_synthetic_device_generator_matches_law_and_oracle(cuda_required, tks, orc):
    """tks_generate_synthetic: download the matrix it made, check the generator law (create_matrices.py)
    and run the oracle on exactly that matrix."""
    rows, cols = 200_000, 1024
    with tks.SpMV(num_cols=cols, k=100) as s:
        for dist, avg in (("gamma", 20), ("uniform", 40)):
            s.generate_synthetic(rows, cols, avg, dist, seed=7)
            ptr, idx, val = s.download_csr()
            deg = np.diff(ptr.astype(np.int64))
            assert deg.min() >= 1
            if dist == "uniform":
                assert deg.min() == avg // 2 and deg.max() == int(avg * 1.5)
                assert abs(deg.mean() - avg) < 0.2
            else:
                assert abs(deg.mean() - 19.5) < 0.2          # E[int(Gamma(3, 20/3))]
            x = np.repeat(np.arange(rows, dtype=np.uint32), deg)
            key = x.astype(np.int64) * cols + idx
            assert np.all(np.diff(key) >= 0), "columns must be sorted inside a row"
            nrm = np.sqrt(np.add.reduceat(val.astype(np.float64) ** 2, ptr[:-1].astype(np.int64)))
            np.testing.assert_allclose(nrm, 1.0, rtol=1e-5)
            vec = make_query(cols, 3)
            s.reset(vec)
            s()
            v, i, c = s.read_result()
            check_against_scores(i, v, c, orc.spmv_f32(x, idx, val, vec, rows), 100)


def test_two_million_rows_vs_float64(cuda_required, tks, orc):
    """Larger than L2: 2M x 1024 gamma-20 generated in HBM, checked against float64 on the host."""
    rows, cols = 2_000_000, 1024
    with tks.SpMV(num_cols=cols, k=100) as s:
        s.generate_synthetic(rows, cols, 20, "gamma", seed=11)
        ptr, idx, val = s.download_csr()
        for qs in (1, 2):
            vec = make_query(cols, qs)
            s.reset(vec)
            s()
            v, i, c = s.read_result()
            fi, fv = orc.f64_topk(ptr, idx, val, vec, 100)
            np.testing.assert_allclose(v, fv, rtol=RTOL)
            assert len(set(i.tolist()) ^ set(fi.tolist())) <= 2


def test_row_shards_merge_on_device_equals_unsharded(cuda_required, tks, orc, cfg1):
    """The multi-GPU path on one device: 3 contiguous row shards (3 handles), their K candidate keys
    concatenated (what the NCCL all-gather delivers) and merged by tks_merge_keys_device."""
    import torch
    x, y, v, ptr = cfg1
    sh = tks.sharding
    vec = make_query(1024, 31)
    k = 100
    idx0, val0, _ = run_engine(tks, ptr, y, v, 10000, 1024, vec, k)
    shards = sh.plan_row_shards_by_nnz(ptr, 3)
    engines, gathered = [], torch.empty(3 * k, dtype=torch.int64, device="cuda")
    for r, (r0, r1) in enumerate(shards):
        p, i, vv = sh.slice_csr(ptr, y, v, r0, r1)
        e = tks.SpMV(p, i, vv, r1 - r0, 1024, vec=vec, k=k, row_offset=r0)
        e()
        kp, n = e.result_keys_device(0)

        class _A:
            __cuda_array_interface__ = {"shape": (k,), "typestr": "<i8", "data": (kp, False), "version": 2}
        gathered[r * k:(r + 1) * k] = torch.as_tensor(_A(), device="cuda")
        engines.append(e)
    torch.cuda.synchronize()
    engines[0].merge_keys_device(gathered.data_ptr(), 3 * k, k)
    val, idx, cnt = engines[0].read_result()
    assert cnt == k
    # same rows; the scores may differ in the last bit because a shard changes which lanes a row's
    # non-zeros fall into (a different fp32 summation order)
    assert np.array_equal(np.sort(idx), np.sort(idx0))
    np.testing.assert_allclose(val, val0, rtol=1e-6)
    for e in engines:
        e.close()


# ---- half-precision value mode (the reference's -a flag, host_spmv_topk_csr_gpu.cu:132-136,151-153) ----------

def test_half_mode_matches_gold_on_half_rounded_inputs(cuda_required, tks, orc, cfg1):
    """TKS_VALUE_FP16: values and query rounded to IEEE half (RNE), exact fp32 products, fp32 row sums."""
    x, y, v, ptr = cfg1
    vh = orc.half_round(v)
    with tks.SpMV(ptr, y, v, 10000, 1024, k=100, half=True) as s:
        assert s.stats().device_bytes < 4.2 * v.size + 8 * 1024            # 2 + 2 + 1/8 bytes per non-zero
        _, _, dval = s.download_csr()
        assert np.array_equal(dval.view(np.uint32), vh.view(np.uint32))     # resident values = RNE halves
        for qs in range(1, 4):
            vec = make_query(1024, qs)
            s.reset(vec)
            s()
            val, idx, cnt = s.read_result()
            yref = orc.spmv_f32(x, y, vh, orc.half_round(vec), 10000)
            check_against_scores(idx, val, cnt, yref, 100)
            gi, gv = orc.gold_topk_f16(x, y, v, vec, 100)
            np.testing.assert_allclose(val, gv, rtol=RTOL)
            # and it stays close to the exact fp32 answer: half inputs carry 11 significant bits
            fi, fv = orc.gold_topk_f32(x, y, v, vec, 100)
            np.testing.assert_allclose(val, fv, rtol=2e-3)
            assert len(set(idx.tolist()) & set(fi.tolist())) >= 90


@pytest.mark.parametrize("k,chunk_nnz", [(1, 256), (100, 128), (500, 4096), (1024, 1024)])
def test_half_mode_k_and_chunk_sweep(cuda_required, tks, orc, cfg1, k, chunk_nnz):
    x, y, v, ptr = cfg1
    vec = make_query(1024, 11)
    idx, val, cnt = run_engine(tks, ptr, y, v, 10000, 1024, vec, k, half=True, chunk_nnz=chunk_nnz)
    check_against_scores(idx, val, cnt, orc.spmv_f32(x, y, orc.half_round(v), orc.half_round(vec), 10000), k)


@pytest.mark.parametrize("tie_higher", [False, True])
def test_half_mode_exact_ties_and_empty_rows(cuda_required, tks, orc, tie_higher):
    """Powers of two are exact in half: the result must be bit-exact including the tie order."""
    rng = np.random.default_rng(99)
    rows, k = 4000, 100
    ptr, x, col, val, vec = exact_matrix(rows, 256, rng, 9, 0.15)
    yref = orc.spmv_f32(x, col, val, vec, rows)
    nonempty = np.diff(ptr.astype(np.int64)) > 0
    with tks.SpMV(ptr, col, val, rows, 256, vec=vec, k=k, tie_higher=tie_higher, chunk_nnz=256, half=True) as s:
        s()
        v, i, c = s.read_result()
    cand = np.nonzero(nonempty)[0]
    order = cand[np.lexsort((cand if not tie_higher else -cand, -yref[cand].astype(np.float64)))][:k]
    assert c == min(k, cand.size)
    assert np.array_equal(i[:c], order.astype(np.uint32))
    assert np.array_equal(v[:c].view(np.uint32), yref[order].view(np.uint32))


def test_half_mode_batched_queries(cuda_required, tks, orc, cfg1):
    """Batched pass over half values: sequential fp32 sums of exact products -> bit-identical to the gold's arithmetic."""
    x, y, v, ptr = cfg1
    vecs = np.stack([make_query(1024, 40 + q) for q in range(5)])
    with tks.SpMV(ptr, y, v, 10000, 1024, k=50, max_batch=8, half=True) as s:
        s.reset(vecs)
        s()
        for q in range(5):
            val, idx, cnt = s.read_result(q)
            yref = orc.spmv_f32(x, y, orc.half_round(v), orc.half_round(vecs[q]), 10000)
            check_against_scores(idx, val, cnt, yref, 50)
            assert np.array_equal(val.view(np.uint32), yref[idx].view(np.uint32))


def test_half_mode_rejected_in_fixed_mode(tks):
    cfg = tks.capi.default_config(mode=tks.capi.MODE_FIXED_BSCSR, value_type=tks.capi.VALUE_FP16)
    import ctypes as C
    h = C.c_void_p()
    assert tks.capi.lib().tks_create(C.byref(cfg), C.byref(h)) == tks.capi.TKS_EINVAL


def test_profile_switch_and_direct_host_results(cuda_required, tks, orc, cfg1):
    """tks_run writes indices / scores / count straight into the pinned host block; tks_run_async + read_result fetches
    them with a copy; with kernel profiling switched on at run time the kernels run without launch overlap and the
    dominant kernel's time is reported.  All three give the same answer."""
    x, y, v, ptr = cfg1
    vec = make_query(1024, 21)
    with tks.SpMV(ptr, y, v, 10000, 1024, vec=vec, k=100) as s:
        s()
        a = s.read_result()
        assert s.stats().last_main_kernel_ms == 0.0
        s.run_async(100)
        b = s.read_result()
        s.set_profile_kernels(True)
        s()
        c = s.read_result()
        assert 0.0 < s.stats().last_main_kernel_ms < s.stats().last_kernel_ms
        s.set_profile_kernels(False)
        s()
        d = s.read_result()
    for r in (b, c, d):
        assert r[2] == a[2] and np.array_equal(r[1], a[1]) and np.array_equal(r[0].view(np.uint32), a[0].view(np.uint32))
    check_against_scores(a[1], a[0], a[2], orc.spmv_f32(x, y, v, vec, 10000), 100)


# ---- bfloat16 value mode (not a mode of the reference) ---------------------------------------------

def test_bf16_mode_matches_gold_on_bf16_rounded_inputs(cuda_required, tks, orc, cfg1):
    x, y, v, ptr = cfg1
    vb = orc.bf16_round(v)
    with tks.SpMV(ptr, y, v, 10000, 1024, k=100, bf16=True) as s:
        assert s.stats().device_bytes < 4.2 * v.size + 8 * 1024
        _, _, dval = s.download_csr()
        assert np.array_equal(dval.view(np.uint32), vb.view(np.uint32))       # resident values = RNE bfloat16
        for qs in range(1, 4):
            vec = make_query(1024, qs)
            s.reset(vec)
            s()
            val, idx, cnt = s.read_result()
            yref = orc.spmv_f32(x, y, vb, orc.bf16_round(vec), 10000)
            check_against_scores(idx, val, cnt, yref, 100)
            gi, gv = orc.gold_topk_bf16(x, y, v, vec, 100)
            np.testing.assert_allclose(val, gv, rtol=RTOL)
            fi, fv = orc.gold_topk_f32(x, y, v, vec, 100)                      # 8 significant bits: close, not equal
            np.testing.assert_allclose(val, fv, rtol=2e-2)
            assert len(set(idx.tolist()) & set(fi.tolist())) >= 80


@pytest.mark.parametrize("k,chunk_nnz,tie_higher", [(1, 256, False), (100, 128, True), (1024, 4096, False)])
def test_bf16_mode_k_chunks_and_exact_ties(cuda_required, tks, orc, cfg1, k, chunk_nnz, tie_higher):
    x, y, v, ptr = cfg1
    vec = make_query(1024, 12)
    idx, val, cnt = run_engine(tks, ptr, y, v, 10000, 1024, vec, k, bf16=True, chunk_nnz=chunk_nnz)
    check_against_scores(idx, val, cnt, orc.spmv_f32(x, y, orc.bf16_round(v), orc.bf16_round(vec), 10000), k)
    # powers of two are exact in bfloat16: bit-exact incl. the tie order
    rng = np.random.default_rng(k)
    eptr, ex, ecol, eval_, evec = exact_matrix(3000, 256, rng, 7, 0.1)
    yref = orc.spmv_f32(ex, ecol, eval_, evec, 3000)
    with tks.SpMV(eptr, ecol, eval_, 3000, 256, vec=evec, k=64, tie_higher=tie_higher, chunk_nnz=256, bf16=True) as s:
        s()
        v2, i2, c2 = s.read_result()
    cand = np.nonzero(np.diff(eptr.astype(np.int64)) > 0)[0]
    order = cand[np.lexsort((cand if not tie_higher else -cand, -yref[cand].astype(np.float64)))][:64]
    assert c2 == 64 and np.array_equal(i2, order.astype(np.uint32))
    assert np.array_equal(v2.view(np.uint32), yref[order].view(np.uint32))


def test_bf16_mode_batched_queries(cuda_required, tks, orc, cfg1):
    x, y, v, ptr = cfg1
    vecs = np.stack([make_query(1024, 60 + q) for q in range(4)])
    with tks.SpMV(ptr, y, v, 10000, 1024, k=50, max_batch=8, bf16=True) as s:
        s.reset(vecs)
        s()
        for q in range(4):
            val, idx, cnt = s.read_result(q)
            yref = orc.spmv_f32(x, y, orc.bf16_round(v), orc.bf16_round(vecs[q]), 10000)
            check_against_scores(idx, val, cnt, yref, 50)
            assert np.array_equal(val.view(np.uint32), yref[idx].view(np.uint32))


@pytest.mark.parametrize("half,batch", [(False, 1), (True, 1), (False, 64)])
def test_work_units_are_a_whole_number_per_resident_warp(cuda_required, tks, half, batch):
    """Default work-unit sizing from 32M non-zeros on: at most m units per resident warp (stream) for a
    whole number m, units of >= 12 000 non-zeros (24 000 with 16-bit values, ~4096 per stream of a batched handle),
    so that no warp is handed one unit more than the others.  Smaller matrices keep 4096-non-zero units."""
    import torch
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    rows = 2_500_000                              # gamma-20: ~4.9e7 non-zeros
    with tks.SpMV(num_cols=1024, k=100, half=half, max_batch=batch) as s:
        s.generate_synthetic(rows, 1024, 20, "gamma", seed=3)
        st = s.stats()
        nnz, unit, units = int(st.nnz), int(st.work_unit_nnz), int(st.work_units)
        assert nnz >= 32 << 20 and units == -(-nnz // unit)
        if batch > 1:
            streams = sms * 24 * 8                # one 768-thread CTA per SM, eight quads per warp
            m = max(1, (-(-nnz // streams)) // 4096)
            assert units <= m * streams and 2048 <= unit <= 8192 + 256
        else:
            # resident warps of the k <= 128 main kernel: 2 x 18 per SM (fp32), 2 x 12 (16-bit values)
            warps = sms * (24 if half else 36)
            tmin = 24000 if half else 12000
            m = max(1, (-(-nnz // warps)) // tmin)
            assert units <= m * warps, (units, m, warps)
            assert unit >= min(tmin, -(-nnz // warps)) and unit % 256 == 0
    with tks.SpMV(num_cols=1024, k=100) as s:
        s.generate_synthetic(200_000, 1024, 20, "gamma", seed=3)
        assert int(s.stats().work_unit_nnz) == 4096
