"""The CUDA path against the committed golden vectors directly (no oracle in between): tests/golden/ holds the
outputs of the REFERENCE's own code (make_golden.py).  Fixed-point: bit-exact kernel result words and merged
list.  Float: scores within 1e-5 relative of the reference gold, same index set except boundary near-ties."""
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"
FIXED = sorted(p.name for p in GOLDEN.glob("fixed_*.npz"))
INPUTS = sorted(p.name for p in GOLDEN.glob("inputs_*.npz"))
RTOL = 1e-5


@pytest.mark.parametrize("name", FIXED)
def test_fixed_engine_equals_reference_outputs(cuda_required, tks, name):
    g = np.load(GOLDEN / name)
    d = np.load(GOLDEN / ("inputs_" + name[len("fixed_"):name.index("_w")] + ".npz"))
    W, Kp, LFR, P, B = (int(g[t]) for t in ("W", "Kp", "LFR", "P", "B"))
    k = 100
    with tks.SpMVFixed(d["x"], d["y"], d["val32"], int(d["rows"]), int(d["cols"]), vec32=d["vec32"], k=k, fixed_width=W,
                       partitions=P, local_k=Kp, limited_finished_rows=LFR) as f:
        assert np.array_equal(f.first_row, g["first_row"])
        f()
        gv, gi = f.read_result()
        iw, vw = f.read_partition_results()
    assert np.array_equal(iw[:, :, :B], g["idx_words"]), "kernel index words differ from the reference kernel's"
    assert np.array_equal(vw[:, :, :B], g["val_words"]), "kernel value words differ from the reference kernel's"
    n = min(k, g["merged_idx"].size)
    assert np.array_equal(gi, g["merged_idx"][:n]) and np.array_equal(gv, g["merged_val"][:n])


@pytest.mark.parametrize("name", INPUTS)
@pytest.mark.parametrize("k", [1, 8, 100])
def test_float_engine_matches_reference_gold(cuda_required, tks, gen, name, k):
    d = np.load(GOLDEN / name)
    rows, cols = int(d["rows"]), int(d["cols"])
    ptr = gen.csr_from_coo(d["x"], rows)
    with tks.SpMV(ptr, d["y"], d["v"].astype(np.float32), rows, cols, vec=d["vec"], k=k, tie_higher=True) as s:
        s()
        val, idx, cnt = s.read_result()
    ri, rv = d[f"gold_k{k}_idx"], d[f"gold_k{k}_val"]
    assert cnt == k
    np.testing.assert_allclose(val, rv, rtol=RTOL, atol=1e-7)
    # index sets: rows may only differ where their scores are within tolerance of the k-th score
    diff = set(idx.tolist()) ^ set(ri.tolist())
    kth = rv[-1]
    for r in diff:
        s_mine = val[idx == r][0] if r in set(idx.tolist()) else rv[ri == r][0]
        assert abs(s_mine - kth) <= RTOL * abs(kth) + 1e-7, f"row {r} differs and is not a boundary tie"
