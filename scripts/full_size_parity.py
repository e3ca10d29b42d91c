"""Parity at BASELINE's FULL sizes (the -m gpu tests stop at sizes the oracle finishes in seconds).

cfg3: the 10M x 1024 gamma-20 matrix in 20-bit BS-CSR, 32 partitions x LFR 4 x local K 8 -- the engine's raw result
words (every slot of every partition) and the merged list against the oracle's literal sequential kernel, bit for
bit, for the reference semantics and the drift-free mode, host-packed and device-packed.
cfg2: the same matrix in fp32 -- top-100 against the reference gold restatement, slot by slot.
cfg2h: the same with half-precision values and query (fp32 accumulation) against the gold on half-rounded inputs.

    python scripts/full_size_parity.py [--rows 10000000] > gpurun_out/full_size_parity.json

Test infrastructure (it imports oracle/); about two minutes of host time at the full size."""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
import oracle  # noqa: E402
from _pkg import pkg  # noqa: E402


def query(cols, seed):
    rng = np.random.default_rng(seed)
    v = rng.random(cols)
    return (v / np.linalg.norm(v)).astype(np.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10_000_000)
    args = ap.parse_args()
    tks = pkg()
    rows, cols, k = args.rows, 1024, 100
    out = {"rows": rows, "cols": cols, "k": k}
    src = tks.SpMV(num_cols=cols, k=k)
    src.generate_synthetic(rows, cols, 20, "gamma", seed=0)
    ptr, idx, val = src.download_csr()
    x = np.repeat(np.arange(rows, dtype=np.uint32), np.diff(ptr.astype(np.int64)))
    out["nnz"] = int(x.size)
    vec = query(cols, 1)

    # cfg2: fp32 against the gold, slot by slot
    t0 = time.time()
    gi, gv = oracle.gold_topk_f32(x, idx, val, vec, k)
    out["gold_f32_s"] = round(time.time() - t0, 1)
    src.reset(vec); src()
    ev, ei, cnt = src.read_result()
    out["cfg2"] = {"count": int(cnt), "same_index_set": bool(set(ei.tolist()) == set(gi.tolist())),
                   "max_rel_score_diff": float(np.max(np.abs(np.sort(ev)[::-1] - np.sort(gv)[::-1]) / np.sort(gv)[::-1]))}
    # cfg2h: half-precision values and query (the reference's -a mode), fp32 accumulation, against the gold on
    # half-rounded inputs; the engine is fed the same fp32 CSR and rounds it itself
    t0 = time.time()
    hi, hv = oracle.gold_topk_f16(x, idx, val, vec, k)
    out["gold_f16_s"] = round(time.time() - t0, 1)
    with tks.SpMV(ptr, idx, val, rows, cols, vec=vec, k=k, half=True) as eng:
        eng()
        hev, hei, hcnt = eng.read_result()
    out["cfg2h"] = {"count": int(hcnt), "same_index_set": bool(set(hei.tolist()) == set(hi.tolist())),
                    "max_rel_score_diff": float(np.max(np.abs(np.sort(hev)[::-1] - np.sort(hv)[::-1]) / np.sort(hv)[::-1]))}
    src.close()

    # cfg3: fixed point, bit for bit
    val32 = oracle.fx32_from_double(val.astype(np.float64))
    vec32 = oracle.query_fx32_from_f32(vec)
    t0 = time.time()
    packed = oracle.pack_bscsr(x, idx, val32, rows, 32, 20)
    out["oracle_pack_s"] = round(time.time() - t0, 1)
    res = {}
    for drift_free in (False, True):
        t0 = time.time()
        iw, vw = oracle.bscsr_kernel(packed, vec32, 8, 4, drift_free)
        ri, rv = oracle.read_result(iw, vw, packed["first_row"], packed["B"])
        t_or = time.time() - t0
        for device_pack in (False, True):
            with tks.SpMVFixed(x, idx, val32, rows, cols, vec32=vec32, k=k, drift_free=drift_free,
                               device_pack=device_pack) as eng:
                eng()
                fv, fi = eng.read_result()
                eiw, evw = eng.read_partition_results()
            res[f"drift_free={drift_free},device_pack={device_pack}"] = {
                "result_words_identical": bool(np.array_equal(eiw, iw) and np.array_equal(evw, vw)),
                "merged_list_identical": bool(np.array_equal(fi, ri[:k]) and np.array_equal(fv, rv[:k])),
                "oracle_kernel_s": round(t_or, 1)}
    out["cfg3"] = res
    out["all_ok"] = bool(out["cfg2"]["same_index_set"] and out["cfg2"]["max_rel_score_diff"] < 1e-5 and
                         out["cfg2h"]["same_index_set"] and out["cfg2h"]["max_rel_score_diff"] < 1e-5 and
                         all(r["result_words_identical"] and r["merged_list_identical"] for r in res.values()))
    print(json.dumps(out, indent=1))
    sys.exit(0 if out["all_ok"] else 1)


if __name__ == "__main__":
    main()
