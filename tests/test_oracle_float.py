"""Oracle (float path) pinned against the reference's own gold compiled here (oracle/_ref) and
against committed golden vectors.  CPU only."""
import numpy as np
import pytest

from conftest import make_query


def small_matrix(gen, rows=2000, cols=1024, deg=20, dist="gamma", seed=0):
    x, y, v = gen.create_sparse_matrix(rows, cols, deg, dist, seed=seed)
    return x, y, v.astype(np.float32)


@pytest.mark.parametrize("dist,k", [("gamma", 100), ("uniform", 8), ("gamma", 1), ("uniform", 33)])
def test_gold_restatement_equals_reference_gold(orc, gen, dist, k):
    if orc.ref_gold() is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this box)")
    x, y, v = small_matrix(gen, dist=dist, seed=3)
    for qs in range(3):
        vec = make_query(1024, qs + 1)
        i1, v1 = orc.gold_topk_f32(x, y, v, vec, k)
        i2, v2 = orc.ref_gold_topk_f32(x, y, v, vec, k)
        assert np.array_equal(i1, i2)
        assert np.array_equal(v1.view(np.uint32), v2.view(np.uint32))   # bit-exact


def test_gold_unsorted_slots_equal_reference(orc, gen):
    if orc.ref_gold() is None:
        pytest.skip("oracle/_ref not built")
    x, y, v = small_matrix(gen, rows=500, seed=5)
    vec = make_query(1024, 9)
    i1, v1 = orc.gold_topk_f32(x, y, v, vec, 16, sort=False)
    i2, v2 = orc.ref_gold_topk_f32(x, y, v, vec, 16, sort=False)
    assert np.array_equal(i1, i2) and np.array_equal(v1.view(np.uint32), v2.view(np.uint32))


def test_gold_with_ties_equals_reference(orc):
    if orc.ref_gold() is None:
        pytest.skip("oracle/_ref not built")
    # 64 identical rows -> every score ties; exercises `>=` replacement and the higher-index-first sort
    rows = np.repeat(np.arange(64, dtype=np.uint32), 3)
    cols = np.tile(np.array([1, 5, 9], np.uint32), 64)
    vals = np.tile(np.array([0.5, 0.25, 0.125], np.float32), 64)
    vec = make_query(16, 2)
    for k in (1, 4, 8, 64):
        i1, v1 = orc.gold_topk_f32(rows, cols, vals, vec, k)
        i2, v2 = orc.ref_gold_topk_f32(rows, cols, vals, vec, k)
        assert np.array_equal(i1, i2) and np.array_equal(v1, v2)


def test_gold_matches_float64_standin_on_index_set(orc, gen):
    """gold (fp32, sequential) vs the sparse_dot_topn stand-in (float64): same index set unless scores
    tie within 1e-5 relative at the boundary; scores within 1e-5 relative (north-star tolerance)."""
    x, y, v = small_matrix(gen, rows=5000, seed=11)
    ptr = gen.csr_from_coo(x, 5000)
    vec = make_query(1024, 4)
    k = 100
    gi, gv = orc.gold_topk_f32(x, y, v, vec, k)
    fi, fv = orc.f64_topk(ptr, y, v, vec, k)
    np.testing.assert_allclose(np.sort(gv)[::-1], fv, rtol=1e-5)
    diff = set(gi.tolist()) ^ set(fi.tolist())
    if diff:
        kth = fv[-1]
        y64 = orc.spmv_f32(x, y, v, vec, 5000)
        assert all(abs(y64[i] - kth) <= 1e-5 * abs(kth) for i in diff)


def test_sort_tuples_order(orc):
    idx = np.array([3, 7, 1, 9, 4], np.uint32)
    val = np.array([0.5, 0.9, 0.5, 0.1, 0.9], np.float32)
    orc.lib().orc_sort_tuples_f32(5, idx, val)
    assert idx.tolist() == [7, 4, 3, 1, 9]          # equal values: higher index first
    assert val.tolist() == pytest.approx([0.9, 0.9, 0.5, 0.5, 0.1])


def test_half_and_bf16_rounding_helpers_match_an_independent_conversion(orc):
    """oracle.half_round / bf16_round state what the engine's 16-bit value modes do to every matrix value and to the
    query (round to nearest even); torch's CPU conversions are an independent implementation of the same rounding."""
    import torch
    rng = np.random.default_rng(0)
    a = np.concatenate([rng.random(200000).astype(np.float32), (rng.random(1000) * 1e-6).astype(np.float32),
                        np.array([0.0, 1.0, 0.5, 1.0 + 2.0 ** -8, 1.0 + 2.0 ** -9, 1.0 + 3 * 2.0 ** -9, 65504.0], np.float32)])
    t = torch.from_numpy(a)
    assert np.array_equal(orc.half_round(a).view(np.uint32), t.to(torch.float16).to(torch.float32).numpy().view(np.uint32))
    assert np.array_equal(orc.bf16_round(a).view(np.uint32), t.to(torch.bfloat16).to(torch.float32).numpy().view(np.uint32))
    # a 16-bit x 16-bit product is exact in fp32 (11 + 11 and 8 + 8 significand bits): the engine's multiply never rounds
    h, b = orc.half_round(a[:50000]), orc.bf16_round(a[:50000])
    for u in (h, b):
        p32 = u * u[::-1]
        assert np.array_equal(p32.astype(np.float64), u.astype(np.float64) * u[::-1].astype(np.float64))
