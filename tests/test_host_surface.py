"""CPU tests of the host surface exported by the C ABI: packet builder, quantisation, MTX loader,
coo2csr, symbol table.  No GPU, no compute calls."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest


def test_library_exports_every_declared_symbol(tks):
    """Every function declared in include/topkspmv.h is exported by libtopkspmv.so."""
    import re
    from pathlib import Path
    hdr = (Path(__file__).resolve().parent.parent / "include" / "topkspmv.h").read_text()
    declared = set(re.findall(r"\b(tks_[A-Za-z0-9_]+)\s*\(", hdr))
    declared -= {"tks_config", "tks_handle", "tks_stats"}
    L = tks.capi.lib()
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, f"not exported: {missing}"
    assert declared == set(tks.capi.SYMBOLS), declared ^ set(tks.capi.SYMBOLS)
    assert L.tks_version() == 1


def test_create_without_gpu_fails_loudly(tks):
    """There is no CPU fallback: on a box without CUDA, tks_create must fail with a clear message."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(tks.capi.TksError, match="no CUDA device|CUDA"):
        tks.SpMV(num_cols=1024)


def test_create_rejects_knob_values_the_kernels_are_not_built_for(tks):
    """Checked before any CUDA call: LIMITED_FINISHED_ROWS above 4 (types.hpp:76's "clean" design LFR = B is not
    instantiated) must fail at create with a message, not at the first run."""
    import ctypes as C
    L = tks.capi.lib()
    for lfr in (0, 5, 15):
        cfg = tks.capi.default_config(mode=tks.capi.MODE_FIXED_BSCSR, limited_finished_rows=lfr)
        h = C.c_void_p()
        assert L.tks_create(C.byref(cfg), C.byref(h)) == tks.capi.TKS_EINVAL
        assert b"limited_finished_rows" in L.tks_last_error(None)


@pytest.mark.parametrize("W", [20, 21, 25, 26, 32])
def test_packet_size_and_quantisation(tks, orc, W):
    assert tks.capi.bscsr_packet_size(W) == orc.packet_size(W) == 511 // (W + 14)
    rng = np.random.default_rng(W)
    v = np.concatenate([rng.random(2000), [0.0, 1.0, 0.5, 1e-10, 0.999999999, 1.5, 1.9999999]])
    a = tks.capi.fixed32_from_double(v)
    assert np.array_equal(a, orc.fx32_from_double(v))
    L = tks.capi.lib()
    b = np.array([L.tks_fixedW_from_fixed32(int(x), W) for x in a], np.uint32)
    assert np.array_equal(b, orc.fxW_from_fx32(a, W))
    ol = orc.lib()
    assert np.array_equal(b, np.array([ol.orc_fxW_from_fx32(int(x), W) for x in a], np.uint32))


@pytest.mark.parametrize("W,P,rows,deg,dist", [(20, 32, 3000, 20, "gamma"), (32, 32, 3000, 20, "uniform"),
                                                (25, 8, 500, 3, "gamma"), (21, 4, 64, 40, "uniform"),
                                                (26, 1, 100, 2, "gamma"), (20, 32, 32, 1, "gamma")])
def test_packer_bit_exact_vs_oracle(tks, orc, gen, W, P, rows, deg, dist):
    """Host packet builder (the product's) == oracle's literal transcription of host_spmv_bscsr.cpp:189-248."""
    x, y, v = gen.create_sparse_matrix(rows, 1024, max(deg, 2), dist, seed=W + P)
    val32 = orc.fx32_from_double(v)
    packets, ppp, first, npp = tks.capi.pack_bscsr(x, y, val32, rows, P, W)
    o = orc.pack_bscsr(x, y, val32, rows, P, W)
    assert np.array_equal(ppp, o["num_packets"])
    assert np.array_equal(first, o["first_row"])
    assert np.array_equal(packets, np.concatenate(o["packets"]))
    assert int(npp.sum()) == x.size


def test_packer_rejects_empty_partition_and_unsorted(tks):
    x = np.array([0, 0, 1, 5], np.uint32)
    y = np.zeros(4, np.uint32)
    v = np.ones(4, np.uint32)
    with pytest.raises(tks.capi.TksError, match="no non-zeros"):
        tks.capi.pack_bscsr(x, y, v, 64, 32, 20)
    with pytest.raises(tks.capi.TksError):
        tks.capi.pack_bscsr(np.array([3, 1, 2], np.uint32), y[:3], v[:3], 4, 1, 20)


def test_mtx_loader_roundtrip_and_vs_reference(tks, orc, gen, tmp_path):
    rows, cols = 400, 512
    x, y, v = gen.create_sparse_matrix(rows, cols, 20, "gamma", seed=2)
    for zero in (False, True):
        p = tmp_path / f"m{int(zero)}.mtx"
        gen.write_mtx(p, x, y, v, rows, cols, zero_indexed=zero)
        r, c, xx, yy, vv = tks.capi.read_mtx(p, zero_indexed=zero)
        assert (r, c) == (rows, cols)
        assert np.array_equal(xx, x) and np.array_equal(yy, y)
        np.testing.assert_allclose(vv, v, rtol=1e-9)          # 10 significant digits in the file
        if orc.ref_gold() is not None:                          # the reference's own readMtx on the same file
            rc, rr, rcols, rx, ry, rv = orc.ref_read_mtx(p, zero_indexed=zero, sort=False)
            assert (rr, rcols) == (rows, cols)
            assert np.array_equal(rx, xx) and np.array_equal(ry, yy)
            assert np.array_equal(rv, vv.astype(np.float32))
    # header is exactly the reference generator's three lines (test_cpu.py parses lines[2])
    lines = open(tmp_path / "m0.mtx").read().split("\n")
    assert lines[0] == "%%MatrixMarket matrix coordinate real general" and lines[1] == "%"
    assert lines[2] == f"{rows} {cols} {len(x)}"


def test_mtx_loader_variants(tks, tmp_path):
    p = tmp_path / "sym.mtx"
    p.write_text("%%MatrixMarket matrix coordinate real symmetric\n% c\n% c2\n3 3 3\n1 1 1.0\n2 1 2.5\n3 2 -1e-1\n")
    r, c, x, y, v = tks.capi.read_mtx(p, sort_tuples=True)
    assert (r, c) == (3, 3)
    assert sorted(zip(x.tolist(), y.tolist(), v.tolist())) == [(0, 0, 1.0), (0, 1, 2.5), (1, 0, 2.5), (1, 2, -0.1), (2, 1, -0.1)]
    q = tmp_path / "pat.mtx"
    q.write_text("%%MatrixMarket matrix coordinate pattern general\n2 2 2\n1 2\n2 1\n")
    r, c, x, y, v = tks.capi.read_mtx(q)
    assert v.tolist() == [1.0, 1.0] and x.tolist() == [0, 1] and y.tolist() == [1, 0]
    with pytest.raises(tks.capi.TksError):
        tks.capi.read_mtx(tmp_path / "missing.mtx")
    bad = tmp_path / "bad.mtx"
    bad.write_text("%%MatrixMarket matrix array real general\n2 2\n1\n2\n3\n4\n")
    with pytest.raises(tks.capi.TksError, match="coordinate"):
        tks.capi.read_mtx(bad)
    short = tmp_path / "short.mtx"
    short.write_text("%%MatrixMarket matrix coordinate real general\n2 2 3\n1 1 1.0\n")
    with pytest.raises(tks.capi.TksError, match="Not enough"):
        tks.capi.read_mtx(short)


def test_coo2csr_vs_reference(tks, orc, gen):
    x, y, v = gen.create_sparse_matrix(300, 64, 6, "uniform", seed=1)
    ptr, idx, val = tks.capi.coo2csr(x, y, v.astype(np.float32), 300, 64)
    assert np.array_equal(ptr.astype(np.uint64), gen.csr_from_coo(x, 300))
    assert np.array_equal(idx, y) and np.array_equal(val, v.astype(np.float32))
    R = orc.ref_gold()
    if R is not None:
        rp = np.zeros(301, np.uint32); ri = np.zeros(x.size, np.uint32); rv = np.zeros(x.size, np.float32)
        R.ref_coo2csr(x, y, v.astype(np.float32), x.size, 300, 64, rp, ri, rv)
        assert np.array_equal(rp, ptr) and np.array_equal(ri, idx) and np.array_equal(rv, val)
    with pytest.raises(tks.capi.TksError):
        tks.capi.coo2csr(np.array([5], np.uint32), np.array([0], np.uint32), np.ones(1, np.float32), 3, 3)


def test_sharding_plans(tks):
    sh = tks.sharding
    ptr = np.concatenate([[0], np.cumsum(np.random.default_rng(0).integers(1, 50, 1000))]).astype(np.uint64)
    for n in (1, 2, 3, 8):
        plan = sh.plan_row_shards_by_nnz(ptr, n)
        assert plan[0][0] == 0 and plan[-1][1] == 1000
        assert all(plan[i][1] == plan[i + 1][0] for i in range(n - 1))
        sizes = [int(ptr[e] - ptr[b]) for b, e in plan]
        assert max(sizes) - min(sizes) <= 100
        even = sh.plan_row_shards_even(1000, n)
        assert sum(e - b for b, e in even) == 1000 and max(e - b for b, e in even) - min(e - b for b, e in even) <= 1
    scores = np.array([0.5, 0.25, 0.5, -1.0, 0.0], np.float32)
    rows = np.array([7, 3, 2, 9, 1], np.uint32)
    for th in (False, True):
        keys = sh.make_keys(scores, rows, th)
        s2, r2 = sh.split_keys(keys, th)
        assert np.array_equal(s2, scores) and np.array_equal(r2, rows)
        top = sh.merge_topk_host([keys[:2], keys[2:]], 3)
        s3, r3 = sh.split_keys(top, th)
        assert s3.tolist() == [0.5, 0.5, 0.25] and r3.tolist() == ([7, 2, 3] if th else [2, 7, 3])


def test_merge_partition_words_equals_reference_read_result(tks, orc, gen):
    """tks_merge_partition_words = read_result + sort_tuples (host_spmv_bscsr.cpp:399-448) over the kernel's result
    words; the oracle's read_result is pinned against the compiled reference (tests/test_oracle_fixed.py)."""
    x, y, v = gen.create_sparse_matrix(10000, 1024, 20, "gamma", seed=0)
    rng = np.random.default_rng(7)
    vec = rng.random(1024); vec = (vec / np.linalg.norm(vec)).astype(np.float32)
    for W, P, Kp, LFR in [(20, 32, 8, 4), (32, 8, 4, 2), (26, 32, 8, 4)]:
        o = orc.bscsr_topk(x, y, v, 10000, vec, P=P, W=W, Kp=Kp, LFR=LFR)
        val, idx = tks.capi.merge_partition_words(o["idx_words"], o["val_words"], o["packed"]["first_row"],
                                                  tks.capi.bscsr_packet_size(W), 4096)
        assert np.array_equal(idx, o["idx"]) and np.array_equal(val, o["val"])
        val, idx = tks.capi.merge_partition_words(o["idx_words"], o["val_words"], o["packed"]["first_row"],
                                                  tks.capi.bscsr_packet_size(W), 10)
        assert np.array_equal(idx, o["idx"][:10]) and np.array_equal(val, o["val"][:10])
    # duplicates of an index across slots: the first insertion wins; zero values are dropped
    iw = np.zeros((2, 2, 16), np.uint32); vw = np.zeros((2, 2, 16), np.uint32)
    iw[0, 0, 0], vw[0, 0, 0] = 5, 100
    iw[0, 1, 1], vw[0, 1, 1] = 5, 900          # same row again, later slot: ignored
    iw[1, 0, 0], vw[1, 0, 0] = 1, 100          # row 1 + first_row 10 = 11: ties with row 5 -> higher index first
    iw[1, 1, 3], vw[1, 1, 3] = 2, 0            # val == 0: not a candidate
    val, idx = tks.capi.merge_partition_words(iw, vw, np.array([0, 10], np.uint32), 15, 8)
    assert idx.tolist() == [11, 5] and val.tolist() == [100, 100]


def test_partition_shards_keep_the_unsharded_boundaries(tks):
    """FPGA mode over several ranks: P / N partitions per rank, rows_per_part of the UNSHARDED matrix (ceil(N / P),
    host_spmv_bscsr.cpp:136); the last shard may be short, every other one is exactly parts_per_rank * rows_per_part."""
    plan = tks.distributed.plan_partition_shards
    rpp, ppr, shards = plan(10_000_000, 32, 8)
    assert (rpp, ppr) == (312500, 4) and shards[0] == (0, 1250000) and shards[-1] == (8750000, 10_000_000)
    rpp, ppr, shards = plan(7777, 8, 2)
    assert (rpp, ppr) == (973, 4) and shards == [(0, 3892), (3892, 7777)]
    rpp, ppr, shards = plan(100, 32, 4)          # more partitions than some shards have rows: clipped, never negative
    assert rpp == 4 and all(0 <= a <= b <= 100 for a, b in shards) and shards[-1][1] == 100
    with pytest.raises(ValueError):
        plan(1000, 32, 3)


def test_exchange_mode_selection_without_peers(tks):
    """One rank (or no NCCL): "auto" quietly uses the all-gather path, an explicit "peer" request is refused loudly."""
    s = tks.ShardedSpMV(engine=None, k=10)
    assert s.exchange_mode == "nccl" and s.world == 1
    assert tks.ShardedSpMV(engine=None, k=10, exchange="none").exchange_mode == "none"
    with pytest.raises(ValueError):
        tks.ShardedSpMV(engine=None, k=10, exchange="peer")


def test_cpu_driver_runs_and_keeps_the_reference_csv_schema(gen, tmp_path):
    """test_cpu.py (BASELINE config 1's driver): same CLI and CSV columns as the reference's (test_cpu.py:30-122)."""
    import pandas as pd
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    x, y, v = gen.create_sparse_matrix(2000, 1024, 20, "gamma", seed=0)
    mtx = tmp_path / gen.matrix_name(2000, 1024, 20, "gamma")
    gen.write_mtx(mtx, x, y, v, 2000, 1024)
    out = tmp_path / "cpu.csv"
    r = subprocess.run([sys.executable, os.path.join(root, "test_cpu.py"), "-i", str(mtx), "-t", "3", "-k", "100", "-o", str(out),
                        "-s", "1"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-1000:]
    df = pd.read_csv(out)
    assert list(df.columns) == ["iter", "rows", "cols", "nnz", "K", "exec_time_ms"] and len(df) == 3
    assert int(df.rows[0]) == 2000 and int(df.cols[0]) == 1024 and int(df.nnz[0]) == x.size and int(df.K[0]) == 100


def test_default_torch_stream_is_named_by_cuda_stream_legacy():
    """The C ABI reads a NULL stream as "the handle's private stream"; torch's default stream is handle 0.  The Python
    layer must hand it over as cudaStreamLegacy (1) so that collectives and copies on torch's stream are ordered behind
    the engine's kernels; any other stream goes through unchanged."""
    from types import SimpleNamespace
    from _pkg import pkg
    dist = pkg().distributed

    def fake(handle):
        return SimpleNamespace(cuda=SimpleNamespace(current_stream=lambda: SimpleNamespace(cuda_stream=handle)))
    assert dist._current_stream_handle(fake(0)) == 1
    assert dist._current_stream_handle(fake(0x7F00DEAD0000)) == 0x7F00DEAD0000
