"""GPU-side BS-CSR packer (csrc/bscsr_pack.cuh) against the host packer + upload.

The host path (tks_pack_bscsr -> tks_upload_bscsr) is itself pinned, bit for bit, to the reference's
packet_coo / packet_coo_partition (tests/test_oracle_fixed.py, tests/test_golden.py).  The device path must
leave byte-identical resident state -- packets (verbatim or BSX re-encoded), chunk tables with the tabulated
row counters and carry look-backs, sample tables, partition first rows -- and therefore bit-identical results."""
import ctypes as C

import numpy as np
import pytest

from conftest import make_query

pytestmark = pytest.mark.gpu


def both_paths(tks, orc, x, y, v, rows, cols, W=20, P=32, Kp=8, LFR=4, drift_free=False, k=100):
    val32 = orc.fx32_from_double(v)
    vec32 = orc.query_fx32_from_f32(make_query(cols, 5))
    kw = dict(vec32=vec32, k=k, fixed_width=W, partitions=P, local_k=Kp, limited_finished_rows=LFR, drift_free=drift_free)
    with tks.SpMVFixed(x, y, val32, rows, cols, **kw) as host, \
            tks.SpMVFixed(x, y, val32, rows, cols, device_pack=True, **kw) as dev:
        dh, dd = host.state_digest(), dev.state_digest()
        names = ["packets", "chunk_first", "chunk_count", "chunk_local0", "chunk_row_in", "chunk_lookback", "chunk_part",
                 "part_chunk_begin", "s_first", "s_count", "s_local0", "s_lookback", "s_part", "s_part_begin",
                 "sample_end", "first_row+counts"]
        diff = [n for n, a, b in zip(names, dh, dd) if a != b]
        assert not diff, f"device-packed state differs from host-packed state in: {diff}"
        host(); dev()
        hv, hi = host.read_result()
        dv, di = dev.read_result()
        assert np.array_equal(hi, di) and np.array_equal(hv, dv)
        hw, dw = host.read_partition_results(), dev.read_partition_results()
        assert np.array_equal(hw[0], dw[0]) and np.array_equal(hw[1], dw[1])
        sh, sd = host.stats(), dev.stats()
        assert (sh.packets, sh.nnz, sh.algorithmic_bytes) == (sd.packets, sd.nnz, sd.algorithmic_bytes)
    return hi, hv


@pytest.mark.parametrize("W", [20, 21, 25, 26, 32])
def test_device_packer_all_widths_cfg1(cuda_required, tks, orc, gen, W):
    x, y, v = gen.create_sparse_matrix(10000, 1024, 20, "gamma", seed=0)
    both_paths(tks, orc, x, y, v, 10000, 1024, W=W)


@pytest.mark.parametrize("deg,dist,LFR,P", [(2, "gamma", 4, 32), (3, "gamma", 2, 8), (6, "uniform", 3, 32),
                                            (40, "uniform", 4, 4), (20, "gamma", 1, 64)])
def test_device_packer_row_shapes_and_knobs(cuda_required, tks, orc, gen, deg, dist, LFR, P):
    """Short rows (packets with more than LFR segments: the drifting row counter must be tabulated identically),
    long rows (packets that only pass the carry through: look-back runs), few / many partitions."""
    x, y, v = gen.create_sparse_matrix(30000, 1024, deg, dist, seed=deg)
    both_paths(tks, orc, x, y, v, 30000, 1024, LFR=LFR, P=P)


@pytest.mark.parametrize("deg,LFR,W", [(2, 4, 20), (3, 2, 20), (4, 3, 21), (20, 4, 20)])
def test_device_packer_drift_free(cuda_required, tks, orc, gen, deg, LFR, W):
    x, y, v = gen.create_sparse_matrix(40000, 1024, deg, "gamma", seed=deg + LFR)
    both_paths(tks, orc, x, y, v, 40000, 1024, W=W, LFR=LFR, drift_free=True)


def test_device_packer_verbatim_narrow(cuda_required, tks, orc, gen, monkeypatch):
    monkeypatch.setenv("TKS_BSCSR_VERBATIM", "1")
    x, y, v = gen.create_sparse_matrix(20000, 1024, 5, "gamma", seed=7)
    both_paths(tks, orc, x, y, v, 20000, 1024, W=20)


def test_device_packer_large_enough_for_many_chunks(cuda_required, tks, orc, gen):
    """400k rows: ~1000 chunks, the short tail chunks, sample pieces capped by kBsSamplePackets; result vs the oracle."""
    rows = 400000
    x, y, v = gen.create_sparse_matrix(rows, 1024, 20, "gamma", seed=11)
    hi, hv = both_paths(tks, orc, x, y, v, rows, 1024)
    o = orc.bscsr_topk(x, y, v, rows, make_query(1024, 5))
    n = min(100, o["idx"].size)
    assert np.array_equal(hi, o["idx"][:n]) and np.array_equal(hv, o["val"][:n])


def test_device_packer_rejects_bad_input(cuda_required, tks, orc, gen):
    x, y, v = gen.create_sparse_matrix(5000, 1024, 20, "gamma", seed=1)
    val32 = orc.fx32_from_double(v)
    xs = x.copy(); xs[10], xs[2000] = xs[2000], xs[10]                     # unsorted rows
    with pytest.raises(tks.capi.TksError, match="not sorted"):
        tks.SpMVFixed(xs, y, val32, 5000, 1024, device_pack=True)
    yb = y.copy(); yb[5] = 1024                                            # column out of range
    with pytest.raises(tks.capi.TksError, match="column index"):
        tks.SpMVFixed(x, yb, val32, 5000, 1024, device_pack=True)
    keep = x < 100                                                          # rows 100.. empty -> empty partitions
    with pytest.raises(tks.capi.TksError, match="no non-zeros"):
        tks.SpMVFixed(x[keep], y[keep], val32[keep], 5000, 1024, device_pack=True)
