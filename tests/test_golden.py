"""The oracle against the committed golden vectors (tests/golden/, made by make_golden.py from the REFERENCE's
own code: gold_algorithms.hpp for the float path; host_spmv_bscsr.cpp + the HLS kernel, compiled against
oracle/shim, for the fixed-point path).  CPU only; this is what pins the oracle where /root/reference is absent."""
import hashlib
from pathlib import Path

import numpy as np
import pytest

GOLDEN = Path(__file__).resolve().parent / "golden"
FIXED = sorted(p.name for p in GOLDEN.glob("fixed_*.npz"))
INPUTS = sorted(p.name for p in GOLDEN.glob("inputs_*.npz"))


def load_fixed(name):
    g = np.load(GOLDEN / name)
    matrix = name[len("fixed_"):name.index("_w")]
    d = np.load(GOLDEN / f"inputs_{matrix}.npz")
    return g, d


def test_fixtures_exist():
    assert len(FIXED) >= 20 and len(INPUTS) >= 5


@pytest.mark.parametrize("name", INPUTS)
def test_float_gold_restatement_equals_reference_outputs(orc, name):
    """gold_algorithms.hpp:188-246: same slots (unsorted) and same sort_tuples order, bit for bit."""
    d = np.load(GOLDEN / name)
    v32 = d["v"].astype(np.float32)
    for k in (1, 8, 100):
        ui, uv = orc.gold_topk_f32(d["x"], d["y"], v32, d["vec"], k, sort=False)
        assert np.array_equal(ui, d[f"gold_k{k}_slots_idx"]) and np.array_equal(uv, d[f"gold_k{k}_slots_val"])
        si, sv = orc.gold_topk_f32(d["x"], d["y"], v32, d["vec"], k, sort=True)
        assert np.array_equal(si, d[f"gold_k{k}_idx"]) and np.array_equal(sv, d[f"gold_k{k}_val"])


@pytest.mark.parametrize("name", FIXED)
def test_fixed_oracle_equals_reference_outputs(orc, name):
    """Quantisation, partitioning, packets, packed query, kernel result words, host merge: bit-exact."""
    g, d = load_fixed(name)
    W, Kp, LFR, P, B = (int(g[t]) for t in ("W", "Kp", "LFR", "P", "B"))
    assert orc.packet_size(W) == B
    assert np.array_equal(orc.fx32_from_double(g["probe"]), g["probe_fx32"])
    assert np.array_equal(orc.fxW_from_fx32(g["probe_fx32"], W), g["probe_fxW"])
    assert np.array_equal(orc.fx32_from_double(d["v"]), d["val32"])
    packed = orc.pack_bscsr(d["x"], d["y"], d["val32"], int(d["rows"]), P, W)
    assert np.array_equal(packed["first_row"], g["first_row"])
    assert np.array_equal(packed["last_row"], g["last_row"])
    assert np.array_equal(packed["num_packets"].astype(np.uint32), g["packets_per_part"])
    allp = np.concatenate(packed["packets"], axis=0)
    assert np.array_equal(allp[:64], g["packets_head"])
    assert hashlib.sha256(np.ascontiguousarray(allp).tobytes()).hexdigest() == str(g["packets_sha256"])
    assert np.array_equal(orc.pack_query(d["vec32"], W), g["query_blocks"])
    iw, vw = orc.bscsr_kernel(packed, d["vec32"], Kp, LFR)
    assert np.array_equal(iw[:, :, :B], g["idx_words"]) and np.array_equal(vw[:, :, :B], g["val_words"])
    ri, rv = orc.read_result(iw, vw, packed["first_row"], B)
    assert np.array_equal(ri, g["merged_idx"]) and np.array_equal(rv, g["merged_val"])
    gi, gv = orc.gold_topk_fx32(d["x"], d["y"], d["val32"], d["vec32"], 100)
    assert np.array_equal(gi, g["gold_fx32_idx"]) and np.array_equal(gv, g["gold_fx32_val"])
