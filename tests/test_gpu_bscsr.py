"""GPU parity tests of the FPGA-semantics engine (W-bit fixed point, BS-CSR packets, P partitions x LFR
lanes x local K) through the C ABI.  Bar: BIT-EXACT against the oracle's literal sequential transcription
of the HLS kernel, both the raw per-partition result words (slot positions included) and the merged list."""
import numpy as np
import pytest

from conftest import make_query

pytestmark = pytest.mark.gpu


def run_both(tks, orc, x, y, v, rows, cols, vec, k=100, W=20, P=32, Kp=8, LFR=4, drift_free=False):
    o = orc.bscsr_topk(x, y, v, rows, vec, P=P, W=W, Kp=Kp, LFR=LFR, drift_free=drift_free)
    with tks.SpMVFixed(x, y, o["val32"], rows, cols, vec32=o["vec32"], k=k, fixed_width=W, partitions=P,
                       local_k=Kp, limited_finished_rows=LFR, drift_free=drift_free) as f:
        f()
        gv, gi = f.read_result()
        iw, vw = f.read_partition_results()
        # a second query on the same handle (reset path)
        vec2 = make_query(cols, 4242)
        o2 = orc.bscsr_topk(x, y, v, rows, vec2, P=P, W=W, Kp=Kp, LFR=LFR, drift_free=drift_free)
        f.reset(o2["vec32"])
        f()
        gv2, gi2 = f.read_result()
    assert np.array_equal(iw, o["idx_words"]), "partition index words differ"
    assert np.array_equal(vw, o["val_words"]), "partition value words differ"
    n = min(k, o["idx"].size)
    assert np.array_equal(gi, o["idx"][:n]) and np.array_equal(gv, o["val"][:n])
    n2 = min(k, o2["idx"].size)
    assert np.array_equal(gi2, o2["idx"][:n2]) and np.array_equal(gv2, o2["val"][:n2])
    return o


@pytest.mark.parametrize("W", [20, 21, 25, 26, 32])
def test_cfg1_all_widths(cuda_required, tks, orc, gen, W):
    """BASELINE config 1 matrix through the designs of test_spmv_topk.py:41-47 (20/21/25/26/32 bit)."""
    x, y, v = gen.create_sparse_matrix(10000, 1024, 20, "gamma", seed=0)
    o = run_both(tks, orc, x, y, v, 10000, 1024, make_query(1024, 1), W=W)
    assert o["idx"].size >= 100


@pytest.mark.parametrize("W", [20, 21])
def test_verbatim_packet_path_for_narrow_widths(cuda_required, tks, orc, gen, W, monkeypatch):
    """FIXED_WIDTH <= 22 normally runs on the re-encoded BSX device format; TKS_BSCSR_VERBATIM=1 keeps the
    reference's 512-bit words on the device (the path of the wider formats).  Both must be bit-exact."""
    monkeypatch.setenv("TKS_BSCSR_VERBATIM", "1")
    x, y, v = gen.create_sparse_matrix(60000, 1024, 20, "gamma", seed=3)
    run_both(tks, orc, x, y, v, 60000, 1024, make_query(1024, 12), W=W, P=4)
    x, y, v = gen.create_sparse_matrix(20000, 1024, 3, "gamma", seed=4)     # LFR overflow drift
    run_both(tks, orc, x, y, v, 20000, 1024, make_query(1024, 13), W=W)


@pytest.mark.parametrize("deg,dist", [(2, "gamma"), (4, "gamma"), (6, "uniform"), (40, "uniform")])
def test_row_lengths_incl_lfr_overflow_drift(cuda_required, tks, orc, gen, deg, dist):
    """Short rows put more than LFR segments into a packet: the reference's row counter drifts and partial
    sums are mis-carried; the engine must reproduce exactly that."""
    x, y, v = gen.create_sparse_matrix(20000, 1024, deg, dist, seed=deg)
    run_both(tks, orc, x, y, v, 20000, 1024, make_query(1024, 2))


@pytest.mark.parametrize("deg,dist,LFR,W", [(2, "gamma", 4, 20), (3, "gamma", 2, 20), (4, "gamma", 3, 21),
                                            (20, "gamma", 4, 20), (1, "uniform", 4, 20)])
def test_drift_free_mode_bit_exact_vs_its_oracle(cuda_required, tks, orc, gen, deg, dist, LFR, W):
    """fixed_drift_free (the repair of the reference's row-counter drift): bit-exact against the
    oracle's statement of the repaired rules, including inputs where almost every packet overflows LFR."""
    x, y, v = gen.create_sparse_matrix(40000, 1024, max(deg, 2), dist, seed=deg + LFR)
    if deg == 1:                                        # one non-zero per row: 15 rows finish in every packet
        x = np.arange(x.size, dtype=np.uint32)
    run_both(tks, orc, x, y, v, int(x.max()) + 1, 1024, make_query(1024, 14), W=W, LFR=LFR, drift_free=True)


def test_drift_free_restores_recall_on_gamma_rows(cuda_required, tks, orc, gen):
    """400k gamma-20 rows: the reference semantics report shifted row indices after the first packet with more than
    LFR segments in a partition (precision collapses); drift-free mode reports the true rows."""
    rows = 400000
    x, y, v = gen.create_sparse_matrix(rows, 1024, 20, "gamma", seed=0)
    vec = make_query(1024, 1)
    yref = orc.spmv_f32(x, y, v.astype(np.float32), vec, rows)
    exact = set(np.argsort(-yref.astype(np.float64), kind="stable")[:100].tolist())
    val32, vec32 = orc.fx32_from_double(v), orc.query_fx32_from_f32(vec)
    prec = {}
    for df in (False, True):
        with tks.SpMVFixed(x, y, val32, rows, 1024, vec32=vec32, k=100, drift_free=df) as f:
            f()
            _, idx = f.read_result()
        prec[df] = len(exact & set(idx.tolist())) / 100
    assert prec[True] >= 0.95, prec
    assert prec[False] < prec[True], prec


def test_drift_free_needs_the_reencoded_format(cuda_required, tks, orc, gen):
    x, y, v = gen.create_sparse_matrix(2000, 1024, 20, "gamma", seed=0)
    val32 = orc.fx32_from_double(v)
    with pytest.raises(tks.capi.TksError, match="fixed_drift_free needs"):
        tks.SpMVFixed(x, y, val32, 2000, 1024, fixed_width=32, drift_free=True)
    with pytest.raises(tks.capi.TksError, match="fixed_drift_free needs"):
        tks.SpMVFixed(x, y, val32, 2000, 1024, fixed_width=20, limited_finished_rows=1, drift_free=True)


@pytest.mark.parametrize("Kp", [1, 2, 4, 8, 16, 32])
def test_local_k_variants_incl_argmin4_typo(cuda_required, tks, orc, gen, Kp):
    x, y, v = gen.create_sparse_matrix(8000, 512, 20, "gamma", seed=Kp)
    run_both(tks, orc, x, y, v, 8000, 512, make_query(512, 3), Kp=Kp, k=50)


@pytest.mark.parametrize("LFR", [1, 2, 3, 4])
def test_limited_finished_rows_variants(cuda_required, tks, orc, gen, LFR):
    x, y, v = gen.create_sparse_matrix(8000, 1024, 12, "gamma", seed=10 + LFR)
    run_both(tks, orc, x, y, v, 8000, 1024, make_query(1024, 5), LFR=LFR)


@pytest.mark.parametrize("P", [1, 2, 7, 32, 64])
def test_partition_counts_and_many_chunks(cuda_required, tks, orc, gen, P):
    """Few partitions => many chunks per partition: exercises the chunk carry look-back, the tabulated row
    counter and the replay filter across chunks."""
    x, y, v = gen.create_sparse_matrix(120000, 1024, 20, "gamma", seed=P)
    run_both(tks, orc, x, y, v, 120000, 1024, make_query(1024, 6), P=P)


def test_massive_ties(cuda_required, tks, orc):
    """Identical rows: every candidate value ties, so the surviving indices depend on the exact slot
    dynamics of the replace-min lists (`>=` replacement, argmin -> highest slot among equal minima)."""
    rows, per = 6000, 5
    x = np.repeat(np.arange(rows, dtype=np.uint32), per)
    y = np.tile(np.array([3, 99, 400, 401, 1000], np.uint32), rows)
    v = np.tile(np.array([0.5, 0.25, 0.125, 0.5, 0.3]), rows)
    # a few distinct rows sprinkled in
    rng = np.random.default_rng(0)
    hot = rng.choice(rows, 40, replace=False)
    for r in hot:
        v[r * per:(r + 1) * per] = rng.random(per)
    for Kp in (4, 8):
        run_both(tks, orc, x, y, v, rows, 1024, make_query(1024, 8), Kp=Kp, P=8)


def test_long_rows_span_many_packets(cuda_required, tks, orc):
    rng = np.random.default_rng(5)
    rows = 640
    deg = rng.integers(1, 400, rows)
    x = np.repeat(np.arange(rows, dtype=np.uint32), deg)
    y = np.sort(rng.integers(0, 1024, x.size)).astype(np.uint32)
    y = rng.integers(0, 1024, x.size).astype(np.uint32)
    v = rng.random(x.size) / 20.0
    run_both(tks, orc, x, y, v, rows, 1024, make_query(1024, 9), P=4, k=30)


def test_tiny_partitions(cuda_required, tks, orc):
    """One row per partition (a single packet each): nothing is ever finished -> empty result."""
    rows = 32
    x = np.arange(rows, dtype=np.uint32)
    y = np.arange(rows, dtype=np.uint32)
    v = np.full(rows, 0.5)
    o = run_both(tks, orc, x, y, v, rows, 64, make_query(64, 1), k=10)
    assert o["idx"].size == 0


def test_errors(cuda_required, tks, orc, gen):
    x, y, v = gen.create_sparse_matrix(2000, 1024, 20, "gamma", seed=0)
    val32 = orc.fx32_from_double(v)
    with pytest.raises(tks.capi.TksError, match="not instantiated|outside"):
        tks.SpMVFixed(x, y, val32, 2000, 1024, fixed_width=23)
    packets, ppp, first, npp = tks.capi.pack_bscsr(x, y, val32, 2000, 32, 20)
    bad = packets.copy()
    bad[5, 0] = (bad[5, 0] & ~np.uint64(0xFF)) | np.uint64(0x12)     # x[0]=2, x[1]=1: decreasing ends
    f = tks.SpMVFixed.__new__(tks.SpMVFixed)
    cfg = tks.capi.default_config(mode=tks.capi.MODE_FIXED_BSCSR, fixed_width=20)
    f._create(cfg); f.num_cols = 1024; f.k = 10
    with pytest.raises(tks.capi.TksError, match="malformed packet"):
        f.upload_packets(bad, ppp, first, npp)
    f.upload_packets(packets, ppp, first, npp)
    with pytest.raises(tks.capi.TksError, match="no query"):
        f()
    f.close()


@pytest.mark.parametrize("P,LFR,Kp", [(4, 4, 8), (2, 2, 4), (8, 3, 16)])
def test_long_candidate_logs_are_pruned_exactly(cuda_required, tks, orc, gen, P, LFR, Kp):
    """Few partitions over many rows: thousands of logged candidates per (partition, lane) list, so the replay kernel's
    parallel pruning (segment thresholds = K-th largest of everything before the segment) does real work; the result
    words must still be the sequential machine's, slot for slot (incl. K = 4 with the reference's argmin_4 typo)."""
    rows = 600000
    x, y, v = gen.create_sparse_matrix(rows, 1024, 20, "gamma", seed=21)
    run_both(tks, orc, x, y, v, rows, 1024, make_query(1024, 9), P=P, LFR=LFR, Kp=Kp)


def test_long_logs_with_massive_ties(cuda_required, tks, orc):
    """One non-zero per pair of rows pattern with values from a tiny set and a constant query: almost every candidate
    ties with the list's minimum -- the `>=` acceptance and the "highest slot among equal minima" rule decide which
    row index survives, and pruning must not change a single slot."""
    rng = np.random.default_rng(5)
    rows, cols = 300000, 64
    deg = rng.integers(1, 4, rows)
    x = np.repeat(np.arange(rows, dtype=np.uint32), deg)
    y = rng.integers(0, cols, x.size).astype(np.uint32)
    v = rng.choice(np.array([0.125, 0.25, 0.5]), x.size)
    vec = np.full(cols, 0.5, np.float32)
    run_both(tks, orc, x, y, v, rows, cols, vec, P=4)


@pytest.mark.parametrize("W,drift_free", [(20, False), (20, True), (32, False)])
def test_pipelined_submits_are_bit_exact(cuda_required, tks, orc, gen, W, drift_free):
    """tks_submit_host / tks_fetch in BS-CSR mode: the sample of query i+1 runs on a second stream beside the stream and
    replay kernels of query i and the replay kernel writes the result words straight to pinned host memory; every
    result must be the oracle's, bit for bit, and the blocking verbs must still work afterwards."""
    rows = 60000
    x, y, v = gen.create_sparse_matrix(rows, 1024, 20, "gamma", seed=9)
    qs = [make_query(1024, 500 + i) for i in range(9)]
    want = [orc.bscsr_topk(x, y, v, rows, q, W=W, drift_free=drift_free) for q in qs]
    with tks.SpMVFixed(x, y, want[0]["val32"], rows, 1024, k=100, fixed_width=W, drift_free=drift_free) as f:
        tickets, got = [], {}
        for i, o in enumerate(want):
            tickets.append(f.submit_host(o["vec32"], 100))
            if i >= 1:
                got[i - 1] = f.fetch(tickets[i - 1])
        got[len(want) - 1] = f.fetch(tickets[-1])
        for i, o in enumerate(want):
            n = min(100, o["idx"].size)
            gv, gi = got[i]
            assert np.array_equal(gi, o["idx"][:n]) and np.array_equal(gv, o["val"][:n]), f"query {i}"
        # a result pushed out of the two slots is refused
        t0 = f.submit_host(want[0]["vec32"], 100)
        f.submit_host(want[1]["vec32"], 100)
        f.submit_host(want[2]["vec32"], 100)
        with pytest.raises(tks.capi.TksError, match="gone"):
            f.fetch(t0)
        f.reset(want[3]["vec32"])
        f()
        gv, gi = f.read_result()
        n = min(100, want[3]["idx"].size)
        assert np.array_equal(gi, want[3]["idx"][:n]) and np.array_equal(gv, want[3]["val"][:n])


def test_pipelined_device_queries_bit_exact(cuda_required, tks, orc, gen):
    import torch
    rows = 60000
    x, y, v = gen.create_sparse_matrix(rows, 1024, 20, "gamma", seed=9)
    qs = [make_query(1024, 800 + i) for i in range(6)]
    want = [orc.bscsr_topk(x, y, v, rows, q) for q in qs]
    dq = torch.from_numpy(np.stack([o["vec32"] for o in want]).view(np.int32)).cuda()
    torch.cuda.synchronize()
    with tks.SpMVFixed(x, y, want[0]["val32"], rows, 1024, k=100) as f:
        for i, o in enumerate(want):
            f.submit(dq[i].data_ptr(), 100, 0, query_ready=True)
            if i % 2 == 1 or i == len(want) - 1:
                gv, gi = f.read_result()
                n = min(100, o["idx"].size)
                assert np.array_equal(gi, o["idx"][:n]) and np.array_equal(gv, o["val"][:n]), f"query {i}"
